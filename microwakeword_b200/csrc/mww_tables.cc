// mww_tables.cc -- host construction of the micro-frontend constant tables (see mww_tables.h).
//
// Configuration = what the reference requests at microwakeword/audio/audio_utils.py:69-81 and what
// pymicro-features hard-wires: 16 kHz, 30 ms / 10 ms, 40 channels over 125..7500 Hz, PCAN
// (strength 0.95, offset 80, gain_bits 21), noise-reduction smoothing_bits 10, log scale_shift 6.
// Float intermediates deliberately use the same float/double mix as the published C library so
// the rounded integer tables come out identical.
#include "mww_tables.h"

#include <math.h>
#include <string.h>

#include <algorithm>

namespace mww {
namespace {

constexpr double kPi = 3.141592653589793238462643383279502884197169399375105820974944;

int16_t q15(double x) { return (int16_t)floor(0.5 + 32767.0 * x); }

float mel_of(float hz) { return (float)(1127.0 * log1p(hz / 700.0)); }

void make_window(HostTables *t) {
    const float arg = (float)(kPi * 2.0 / ((float)kWindow));
    for (int i = 0; i < kWindow; ++i) {
        const float v = (float)(0.5 - (0.5 * cos(arg * (i + 0.5))));
        t->window[i] = (int16_t)floor(v * 4096 + 0.5);
    }
    for (int p = 0; p < 240; ++p) t->win_pairs[p] = pack16(t->window[2 * p], t->window[2 * p + 1]);
}

void make_twiddles(HostTables *t) {
    int16_t re[256], im[256];
    for (int k = 0; k < kNcfft; ++k) {
        const double ph = -2.0 * kPi * k / kNcfft;
        re[k] = q15(cos(ph));
        im[k] = q15(sin(ph));
        t->tw[k] = pack16(re[k], im[k]);
    }
    for (int k = 0; k < kNcfft / 2; ++k) {
        const double ph = -3.14159265358979323846264338327 * ((double)(k + 1) / kNcfft + 0.5);
        t->super_tw[k] = pack16(q15(cos(ph)), q15(sin(ph)));
    }
    // stage 2 of the 4*4*4*4 decimation-in-time plan: butterfly k uses tw[16k], tw[32k], tw[48k]
    for (int k = 1; k <= 3; ++k)
        for (int j = 1; j <= 3; ++j) {
            t->tw2[(k - 1) * 3 + (j - 1)][0] = re[16 * k * j];
            t->tw2[(k - 1) * 3 + (j - 1)][1] = im[16 * k * j];
        }
}

bool make_filterbank(HostTables *t) {
    const int spectrum = kFftSize / 2 + 1;
    const int ranges = kNumChannels + 1;
    const float lo_hz = 125.0f, hi_hz = 7500.0f;
    const float mel_lo = mel_of(lo_hz), mel_hi = mel_of(hi_hz);
    const float spacing = (mel_hi - mel_lo) / ((float)ranges);
    float center[kNumChannels + 1];
    for (int i = 0; i < ranges; ++i) center[i] = mel_lo + (spacing * (i + 1));
    const float hz_per_bin = (float)(0.5 * 16000 / ((float)spectrum - 1));

    for (int b = 0; b < spectrum; ++b) { t->bin_channel[b] = -1; t->bin_weight[b] = 0; t->bin_unweight[b] = 0; }
    t->start_index = (int)(1.5 + lo_hz / hz_per_bin);
    t->end_index = 0;
    int first = t->start_index;
    for (int r = 0; r < ranges; ++r) {
        int b = first;
        while (b < 4096 && mel_of(b * hz_per_bin) <= center[r]) ++b;
        t->chan_start[r] = (int16_t)first;
        const float below = r == 0 ? mel_lo : center[r - 1];
        for (int q = first; q < b && q < spectrum; ++q) {
            const float w = (center[r] - mel_of(q * hz_per_bin)) / (center[r] - below);
            t->bin_channel[q] = (int16_t)r;
            t->bin_weight[q] = (int16_t)floor(w * 4096 + 0.5);
            t->bin_unweight[q] = (int16_t)floor((1.0 - w) * 4096 + 0.5);
        }
        t->end_index = std::max(t->end_index, b);
        first = b;
    }
    t->chan_start[ranges] = (int16_t)first;
    if (t->end_index >= spectrum) return false;

    // Output channel c (0..39) = sum over range c+1 of weight*e + sum over range c of unweight*e:
    // one contiguous span of bins [chan_start[c], chan_start[c+2]) with a merged coefficient row.
    struct Span { int ch, bin0, n; };
    std::vector<Span> spans;
    for (int c = 0; c < kNumChannels; ++c) spans.push_back({c, t->chan_start[c], t->chan_start[c + 2] - t->chan_start[c]});
    std::sort(spans.begin(), spans.end(), [](const Span &a, const Span &b) { return a.n != b.n ? a.n > b.n : a.ch < b.ch; });
    // Every lane runs the same trip count per slot (the loops are unrolled to the slot's longest span), so only the
    // grouping matters: the 16 longest spans share slot 0, the next 16 slot 1, the remaining 8 slot 2.
    std::vector<Span> lane[kFbLanes];
    for (size_t i = 0; i < spans.size(); ++i) lane[i % kFbLanes].push_back(spans[i]);
    for (int s = 0; s < kFbSlots; ++s) {
        t->fb_slot_len[s] = 0;
        for (int l = 0; l < kFbLanes; ++l)
            if ((int)lane[l].size() > s) t->fb_slot_len[s] = std::max(t->fb_slot_len[s], lane[l][s].n);
        t->fb_slot_len[s] = (t->fb_slot_len[s] + 1) & ~1;   // even: energies and coefficients are fetched two per load
    }
    // Start word of every span: bin0 + kEnergyOffset - d with d >= 0 leading zero coefficients, even, n + d within the slot's
    // trip count, and (start / 2) mod 16 distinct over the 16 lanes of a slot (Kuhn's augmenting-path matching of lanes to
    // the 16 bank pairs).  Lanes without a span in a slot take one of the left-over keys (their loads still execute).
    int start_word[kFbLanes][kFbSlots];
    for (int s = 0; s < kFbSlots; ++s) {
        const int L = t->fb_slot_len[s];
        if (L == 0) { for (int l = 0; l < kFbLanes; ++l) start_word[l][s] = 0; continue; }
        int cand[kFbLanes][16];                       // cand[l][key] = start word realising `key` for lane l, or -1
        for (int l = 0; l < kFbLanes; ++l) {
            for (int k = 0; k < 16; ++k) cand[l][k] = -1;
            if ((int)lane[l].size() > s) {
                const Span &sp = lane[l][s];
                for (int d = 0; d + sp.n <= L; ++d) {
                    const int st = sp.bin0 + kEnergyOffset - d;
                    if (st < 0 || (st & 1) || st + L > 272) continue;
                    if (cand[l][(st / 2) % 16] < 0) cand[l][(st / 2) % 16] = st;     // smallest shift wins
                }
            } else {
                for (int k = 0; k < 16; ++k) cand[l][k] = 2 * k;
            }
        }
        int owner[16];
        for (int k = 0; k < 16; ++k) owner[k] = -1;
        struct Aug {
            static bool run(int l, int (*cand)[16], int *owner, bool *seen) {
                for (int k = 0; k < 16; ++k) {
                    if (cand[l][k] < 0 || seen[k]) continue;
                    seen[k] = true;
                    if (owner[k] < 0 || run(owner[k], cand, owner, seen)) { owner[k] = l; return true; }
                }
                return false;
            }
        };
        for (int l = 0; l < kFbLanes; ++l) {
            bool seen[16] = {false};
            if (!Aug::run(l, cand, owner, seen)) return false;      // no conflict-free schedule: the kernel constants need revisiting
        }
        for (int k = 0; k < 16; ++k) start_word[owner[k]][s] = cand[owner[k]][k];
    }
    t->fb_coef.assign(kFbCoefWords, 0);
    int slot_base = 0;
    for (int s = 0; s < kFbSlots; ++s) {
        if (t->fb_slot_len[s] == 0) continue;
        if (s >= 3 || t->fb_slot_len[s] > kFbCoefStride[s]) return false;
        for (int l = 0; l < kFbLanes; ++l) {
            FbSlot &slot = t->fb_slots[l][s];
            slot.coef_off = (int16_t)(slot_base + l * kFbCoefStride[s]);
            slot.word0 = (int16_t)start_word[l][s];
            if ((int)lane[l].size() > s) {
                const Span &sp = lane[l][s];
                slot.ch = (int16_t)sp.ch; slot.n = (int16_t)sp.n;
                for (int j = 0; j < sp.n; ++j) {
                    const int b = sp.bin0 + j;
                    const int w = b + kEnergyOffset - slot.word0;            // position inside the padded span
                    if (w < 0 || w >= t->fb_slot_len[s]) return false;
                    t->fb_coef[slot.coef_off + w] = (b < t->chan_start[sp.ch + 1]) ? t->bin_unweight[b] : t->bin_weight[b];
                }
            } else {
                slot.ch = -1; slot.n = 0;
            }
        }
        slot_base += kFbLanes * kFbCoefStride[s];
    }
    for (int l = 0; l < kFbLanes; ++l)
        for (int s = 0; s < kFbSlots; ++s)
            if (t->fb_slot_len[s] == 0) { FbSlot &slot = t->fb_slots[l][s]; slot.ch = -1; slot.word0 = 0; slot.n = 0; slot.coef_off = 0; }
    return true;
}

int16_t pcan_gain(int input_bits, uint32_t x) {
    const float xf = ((float)x) / ((uint32_t)1 << input_bits);
    const float g = ((uint32_t)1 << 21) * powf(xf + 80.0f, -0.95f);
    if (g > 32767.0f) return 32767;
    return (int16_t)(g + 0.5f);
}

void make_pcan(HostTables *t) {
    memset(t->gain_lut, 0, sizeof t->gain_lut);
    const int input_bits = kSmoothingBits - kLogCorrectionBits;   // 7
    t->gain_lut[0] = pcan_gain(input_bits, 0);
    t->gain_lut[1] = pcan_gain(input_bits, 1);
    for (int k = 2; k <= 32; ++k) {
        const uint32_t x0 = (uint32_t)1 << (k - 1);
        const uint32_t x1 = x0 + (x0 >> 1);
        const uint32_t x2 = (k == 32) ? x0 + (x0 - 1) : 2 * x0;
        const int32_t y0 = pcan_gain(input_bits, x0), y1 = pcan_gain(input_bits, x1), y2 = pcan_gain(input_bits, x2);
        const int32_t a1 = 4 * (y1 - y0) - (y2 - y0);
        const int32_t a2 = (y2 - y0) - a1;
        t->gain_lut[4 * k - 6] = (int16_t)y0;
        t->gain_lut[4 * k - 5] = (int16_t)a1;
        t->gain_lut[4 * k - 4] = (int16_t)a2;
    }
}

void make_log(HostTables *t) {
    memset(t->log_lut, 0, sizeof t->log_lut);
    for (int i = 0; i <= 128; ++i) {
        const double x = i / 128.0;
        t->log_lut[i] = (uint16_t)floor(65536.0 * (log2(1.0 + x) - x) + 0.5);
    }
}

}  // namespace

void build_host_tables(HostTables *t) {
    make_window(t);
    make_twiddles(t);
    t->ok = make_filterbank(t);
    make_pcan(t);
    make_log(t);
}

}  // namespace mww

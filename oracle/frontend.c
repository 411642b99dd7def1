/*
 * oracle/frontend.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the fixed-point "micro-frontend" that the reference
 * calls at microwakeword/audio/audio_utils.py:52-62 (pymicro_features.MicroFrontend
 * .ProcessSamples) and :69-81 (TensorFlow audio_microfrontend op with the same
 * parameters).  The arithmetic lives in an un-vendored, unpinned third-party
 * library (tensorflow/lite/experimental/microfrontend/lib + KissFFT FIXED_POINT=16,
 * as wrapped by PyPI "pymicro-features", setup.py:15) that is absent from
 * /root/reference and from this image, so this file restates the PUBLISHED
 * algorithm from knowledge of the upstream sources.
 *
 *      *** PARITY: pinned to the upstream library's unit-test vectors; not to the reference ***
 * The reference itself holds no golden vectors / tests for this path and the real
 * library cannot run here.  What pins this file: the upstream micro-frontend library's OWN published
 * unit tests (window_test.cc, fft_test.cc, filterbank_test.cc, noise_reduction_test.cc,
 * pcan_gain_control_test.cc, log_scale_test.cc, frontend_test.cc: 1 kHz / 25 ms / 2 channels) -- this
 * code, constructed with that FrontendConfig (mwwo_frontend_create_cfg), reproduces every expected value
 * of all nine stages and the consecutive-frame case exactly (tests/test_upstream_kat.py; the vectors were
 * written down from knowledge of those files, which are not fetchable here, and are cross-checked for
 * mutual consistency), plus closed-form known answers (tests/test_oracle_frontend.py).  The 16 kHz /
 * 40-channel tables are built by the same code.  Every GPU parity claim in this repo is relative to
 * this file.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this code.  The product (microwakeword_b200/) never links it.
 *
 * Configuration restated (audio_utils.py:71-78 + pymicro-features constants):
 *   16 kHz, 30 ms window (480), 10 ms step (160), 40 channels, 125..7500 Hz,
 *   noise reduction: smoothing_bits 10, even 0.025, odd 0.06, min_signal_remaining 0.05
 *   PCAN: on, strength 0.95, offset 80, gain_bits 21;  log: on, scale_shift 6.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "frontend.h"

/* ------------------------------------------------------------------------- */
/* small helpers                                                             */

/* 1-based index of the highest set bit, 0 for x == 0 (upstream bits.h) */
static int msb32(uint32_t x) { return x ? 32 - __builtin_clz(x) : 0; }
static int msb64(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; }

/* ------------------------------------------------------------------------- */
/* KissFFT, FIXED_POINT=16 flavour: int16 samples, Q15 twiddles               */

typedef struct { int16_t r, i; } cpx;

#define KF_FRAC 15
#define KF_SAMP_MAX 32767

/* sround(): add half, arithmetic shift -- result truncated to int16 like upstream */
static int16_t kf_sround(int32_t x) { return (int16_t)((x + (1 << (KF_FRAC - 1))) >> KF_FRAC); }

/* C_FIXDIV(c, div): each component multiplied by SAMP_MAX/div (integer quotient) and rounded */
static void kf_fixdiv(cpx *c, int div) {
    const int32_t k = KF_SAMP_MAX / div;
    c->r = kf_sround((int32_t)c->r * k);
    c->i = kf_sround((int32_t)c->i * k);
}

/* C_MUL: ONE rounding per output component, after the difference / sum of products */
static cpx kf_cmul(cpx a, cpx b) {
    cpx m;
    m.r = kf_sround((int32_t)a.r * b.r - (int32_t)a.i * b.i);
    m.i = kf_sround((int32_t)a.r * b.i + (int32_t)a.i * b.r);
    return m;
}
/* int16 stores wrap, exactly as assigning an int expression to a kiss_fft_scalar does */
static cpx kf_add(cpx a, cpx b) { cpx c = { (int16_t)(a.r + b.r), (int16_t)(a.i + b.i) }; return c; }
static cpx kf_sub(cpx a, cpx b) { cpx c = { (int16_t)(a.r - b.r), (int16_t)(a.i - b.i) }; return c; }

#define NCFFT_MAX 256 /* complex length of the packed real FFT: 256 for the 512-point okay_nabu frame */

struct kf_state {
    int ncfft;                  /* runtime length (fft_size / 2), as kiss_fftr_alloc(nfft) stores it */
    cpx tw[NCFFT_MAX];          /* exp(-2*pi*i*k/ncfft), Q15, floor(.5 + 32767*x) */
    cpx super[NCFFT_MAX / 2];   /* real post-pass twiddles */
    int factors[16];
};

static void kf_factor(int n, int *fac) {
    int p = 4;
    double floor_sqrt = floor(sqrt((double)n));
    do {
        while (n % p) {
            switch (p) {
                case 4: p = 2; break;
                case 2: p = 3; break;
                default: p += 2; break;
            }
            if (p > floor_sqrt) p = n;
        }
        n /= p;
        *fac++ = p;
        *fac++ = n;
    } while (n > 1);
}

static void kf_init(struct kf_state *st, int ncfft) {
    const int NCFFT = ncfft;
    st->ncfft = ncfft;
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    for (int i = 0; i < NCFFT; ++i) {
        double phase = -2.0 * pi * i / NCFFT;
        st->tw[i].r = (int16_t)floor(0.5 + KF_SAMP_MAX * cos(phase));
        st->tw[i].i = (int16_t)floor(0.5 + KF_SAMP_MAX * sin(phase));
    }
    for (int i = 0; i < NCFFT / 2; ++i) {
        double phase = -3.14159265358979323846264338327 * ((double)(i + 1) / NCFFT + 0.5);
        st->super[i].r = (int16_t)floor(0.5 + KF_SAMP_MAX * cos(phase));
        st->super[i].i = (int16_t)floor(0.5 + KF_SAMP_MAX * sin(phase));
    }
    memset(st->factors, 0, sizeof st->factors);
    kf_factor(NCFFT, st->factors);
}

static void kf_bfly2(cpx *F, size_t fstride, const struct kf_state *st, int m) {
    cpx *F2 = F + m;
    const cpx *tw1 = st->tw;
    do {
        kf_fixdiv(F, 2); kf_fixdiv(F2, 2);
        cpx t = kf_cmul(*F2, *tw1);
        tw1 += fstride;
        *F2 = kf_sub(*F, t);
        *F = kf_add(*F, t);
        ++F2; ++F;
    } while (--m);
}

static void kf_bfly4(cpx *F, size_t fstride, const struct kf_state *st, size_t m) {
    const cpx *tw1 = st->tw, *tw2 = st->tw, *tw3 = st->tw;
    const size_t m2 = 2 * m, m3 = 3 * m;
    size_t k = m;
    do {
        kf_fixdiv(&F[0], 4); kf_fixdiv(&F[m], 4); kf_fixdiv(&F[m2], 4); kf_fixdiv(&F[m3], 4);
        cpx s0 = kf_cmul(F[m], *tw1);
        cpx s1 = kf_cmul(F[m2], *tw2);
        cpx s2 = kf_cmul(F[m3], *tw3);
        cpx s5 = kf_sub(F[0], s1);
        F[0] = kf_add(F[0], s1);
        cpx s3 = kf_add(s0, s2);
        cpx s4 = kf_sub(s0, s2);
        F[m2] = kf_sub(F[0], s3);
        tw1 += fstride; tw2 += 2 * fstride; tw3 += 3 * fstride;
        F[0] = kf_add(F[0], s3);
        /* forward transform */
        F[m].r = (int16_t)(s5.r + s4.i);
        F[m].i = (int16_t)(s5.i - s4.r);
        F[m3].r = (int16_t)(s5.r - s4.i);
        F[m3].i = (int16_t)(s5.i + s4.r);
        ++F;
    } while (--k);
}

static void kf_work(cpx *Fout, const cpx *f, size_t fstride, const int *factors, const struct kf_state *st) {
    cpx *beg = Fout;
    const int p = *factors++;
    const int m = *factors++;
    cpx *end = Fout + p * m;
    if (m == 1) {
        do { *Fout = *f; f += fstride; } while (++Fout != end);
    } else {
        do { kf_work(Fout, f, fstride * p, factors, st); f += fstride; } while ((Fout += m) != end);
    }
    Fout = beg;
    switch (p) {
        case 2: kf_bfly2(Fout, fstride, st, m); break;
        case 4: kf_bfly4(Fout, fstride, st, (size_t)m); break;
        default: abort(); /* power-of-two lengths only (256 = 4*4*4*4, 16 = 4*4): radix 3/5/generic never occur */
    }
}

/* kiss_fftr: 2*ncfft real int16 -> ncfft+1 complex int16 (512 -> 257 for okay_nabu) */
static void kf_fftr(const struct kf_state *st, const int16_t *timedata, cpx *freq) {
    const int NCFFT = st->ncfft;
    cpx tmp[NCFFT_MAX];
    kf_work(tmp, (const cpx *)timedata, 1, st->factors, st);

    cpx tdc = tmp[0];
    kf_fixdiv(&tdc, 2);
    freq[0].r = (int16_t)(tdc.r + tdc.i);
    freq[NCFFT].r = (int16_t)(tdc.r - tdc.i);
    freq[NCFFT].i = freq[0].i = 0;

    for (int k = 1; k <= NCFFT / 2; ++k) {
        cpx fpk = tmp[k];
        cpx fpnk = { tmp[NCFFT - k].r, (int16_t)(-tmp[NCFFT - k].i) };
        kf_fixdiv(&fpk, 2);
        kf_fixdiv(&fpnk, 2);
        cpx f1k = kf_add(fpk, fpnk);
        cpx f2k = kf_sub(fpk, fpnk);
        cpx tw = kf_cmul(f2k, st->super[k - 1]);
        freq[k].r = (int16_t)((f1k.r + tw.r) >> 1);
        freq[k].i = (int16_t)((f1k.i + tw.i) >> 1);
        freq[NCFFT - k].r = (int16_t)((f1k.r - tw.r) >> 1);
        freq[NCFFT - k].i = (int16_t)((tw.i - f1k.i) >> 1);
    }
}

/* ------------------------------------------------------------------------- */
/* frontend state                                                            */

#define WIN_MAX 512
#define FFT_MAX 512
#define NCH_MAX MWWO_NUM_CHANNELS
#define WINDOW_BITS 12
#define FB_BITS 12
#define NR_BITS 14
#define PCAN_SNR_BITS 12
#define PCAN_OUT_BITS 6
#define WDF_BITS 32
#define WDF_LUT (4 * WDF_BITS - 3)
#define LOG_SCALE_LOG2 16
#define LOG_SEG_LOG2 7
#define LOG_COEFF 45426

struct mwwo_frontend {
    /* configuration (upstream FrontendConfig): runtime values so that the upstream library's own unit-test
     * configuration (1 kHz, 25 ms window, 2 channels) runs through exactly the code the okay_nabu one does */
    int sample_rate, win, step, fft_n, nch;
    float lower_hz, upper_hz;
    /* window */
    int16_t coef[WIN_MAX];
    int16_t input[WIN_MAX];
    size_t input_used;
    int16_t win_out[WIN_MAX];
    int16_t max_abs;
    /* fft */
    struct kf_state kf;
    int16_t fft_in[FFT_MAX];
    cpx fft_out[FFT_MAX / 2 + 1];
    /* filterbank (kept in "logical" form: per-bin weight + owning channel) */
    int start_index, end_index;
    int16_t bin_channel[FFT_MAX / 2 + 1];   /* channel (0..nch) whose range the bin falls in, -1 outside */
    int16_t bin_weight[FFT_MAX / 2 + 1];
    int16_t bin_unweight[FFT_MAX / 2 + 1];
    int16_t chan_start[NCH_MAX + 2];        /* first bin of each of the nch+1 ranges, + sentinel */
    uint64_t work[NCH_MAX + 1];
    /* noise reduction */
    uint32_t estimate[NCH_MAX];
    uint16_t even_smoothing, odd_smoothing, min_signal_remaining;
    int smoothing_bits;
    /* pcan */
    int16_t gain_lut[WDF_LUT];
    int snr_shift;
    /* log */
    uint16_t log_lut[130];
    int scale_shift;
    int correction_bits;
    /* taps for stage-level tests */
    int last_shift;
    uint32_t last_energy[FFT_MAX / 2 + 1];
    uint64_t last_work[NCH_MAX + 1];
    uint32_t last_sqrt[NCH_MAX];
    uint32_t last_nr[NCH_MAX];
    uint32_t last_pcan[NCH_MAX];
};

static float freq_to_mel(float freq) { return 1127.0 * log1p(freq / 700.0); }

static void build_window(struct mwwo_frontend *s) {
    const int WIN = s->win;
    const float arg = M_PI * 2.0 / ((float)WIN);
    for (int i = 0; i < WIN; ++i) {
        float v = 0.5 - (0.5 * cos(arg * (i + 0.5)));
        s->coef[i] = (int16_t)floor(v * (1 << WINDOW_BITS) + 0.5);
    }
}

static int build_filterbank(struct mwwo_frontend *s) {
    const int nch1 = s->nch + 1;
    const int spectrum = s->fft_n / 2 + 1;
    float center[NCH_MAX + 1];
    const float lower = s->lower_hz, upper = s->upper_hz;
    const float mel_low = freq_to_mel(lower);
    const float mel_hi = freq_to_mel(upper);
    const float mel_span = mel_hi - mel_low;
    const float mel_spacing = mel_span / ((float)nch1);
    for (int i = 0; i < nch1; ++i) center[i] = mel_low + (mel_spacing * (i + 1));

    const float hz_per_sbin = 0.5 * s->sample_rate / ((float)spectrum - 1);
    s->start_index = 1.5 + lower / hz_per_sbin;
    s->end_index = 0;
    for (int b = 0; b < spectrum; ++b) { s->bin_channel[b] = -1; s->bin_weight[b] = s->bin_unweight[b] = 0; }

    int freq_start = s->start_index;
    for (int ch = 0; ch < nch1; ++ch) {
        int f = freq_start;
        while (f < spectrum && freq_to_mel(f * hz_per_sbin) <= center[ch]) ++f;
        s->chan_start[ch] = (int16_t)freq_start;
        const float denom = (ch == 0) ? mel_low : center[ch - 1];
        for (int b = freq_start; b < f; ++b) {
            const float w = (center[ch] - freq_to_mel(b * hz_per_sbin)) / (center[ch] - denom);
            s->bin_channel[b] = (int16_t)ch;
            s->bin_weight[b] = (int16_t)floor(w * (1 << FB_BITS) + 0.5);
            s->bin_unweight[b] = (int16_t)floor((1.0 - w) * (1 << FB_BITS) + 0.5);
        }
        if (f > s->end_index) s->end_index = f;
        freq_start = f;
    }
    s->chan_start[nch1] = (int16_t)freq_start;
    return s->end_index < spectrum;
}

static int16_t pcan_gain_fn(int32_t input_bits, uint32_t x) {
    const float strength = 0.95f, offset = 80.0f;
    const int gain_bits = 21;
    const float xf = ((float)x) / ((uint32_t)1 << input_bits);
    const float g = ((uint32_t)1 << gain_bits) * powf(xf + offset, -strength);
    if (g > 32767.0f) return 32767;
    return (int16_t)(g + 0.5f);
}

static void build_pcan(struct mwwo_frontend *s) {
    const int input_correction_bits = msb32((uint32_t)s->fft_n) - 1 - (FB_BITS / 2);   /* 3 for 512 */
    s->snr_shift = 21 - input_correction_bits - PCAN_SNR_BITS;            /* 6 */
    const int32_t input_bits = s->smoothing_bits - input_correction_bits; /* 7 */
    s->gain_lut[0] = pcan_gain_fn(input_bits, 0);
    s->gain_lut[1] = pcan_gain_fn(input_bits, 1);
    for (int interval = 2; interval <= WDF_BITS; ++interval) {
        const uint32_t x0 = (uint32_t)1 << (interval - 1);
        const uint32_t x1 = x0 + (x0 >> 1);
        const uint32_t x2 = (interval == WDF_BITS) ? x0 + (x0 - 1) : 2 * x0;
        const int16_t y0 = pcan_gain_fn(input_bits, x0);
        const int16_t y1 = pcan_gain_fn(input_bits, x1);
        const int16_t y2 = pcan_gain_fn(input_bits, x2);
        const int32_t d1 = (int32_t)y1 - y0, d2 = (int32_t)y2 - y0;
        const int32_t a1 = 4 * d1 - d2, a2 = d2 - a1;
        int16_t *p = s->gain_lut + 4 * interval - 6;
        p[0] = y0; p[1] = (int16_t)a1; p[2] = (int16_t)a2;
    }
    s->correction_bits = input_correction_bits;
}

static void build_log_lut(struct mwwo_frontend *s) {
    /* closed form of upstream's constant table: round(2^16 * (log2(1 + i/128) - i/128)) */
    for (int i = 0; i <= 128; ++i) {
        double x = (double)i / 128.0;
        s->log_lut[i] = (uint16_t)floor(65536.0 * (log2(1.0 + x) - x) + 0.5);
    }
    s->log_lut[129] = 0;
}

mwwo_frontend *mwwo_frontend_create(void) {
    /* audio_utils.py:71-78: 16 kHz, 30 ms window, 10 ms step, 40 channels, 125..7500 Hz */
    return mwwo_frontend_create_cfg(16000, 30, 10, MWWO_NUM_CHANNELS, 125.0f, 7500.0f);
}

mwwo_frontend *mwwo_frontend_create_cfg(int sample_rate, int window_ms, int step_ms, int num_channels,
                                        float lower_hz, float upper_hz) {
    mwwo_frontend *s = (mwwo_frontend *)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->sample_rate = sample_rate;
    s->win = window_ms * sample_rate / 1000;      /* upstream WindowPopulateState */
    s->step = step_ms * sample_rate / 1000;
    s->nch = num_channels;
    s->lower_hz = lower_hz; s->upper_hz = upper_hz;
    s->fft_n = 1;
    while (s->fft_n < s->win) s->fft_n <<= 1;     /* upstream FftPopulateState: next power of two */
    if (s->win < 2 || s->win > WIN_MAX || s->step < 1 || s->step > s->win || s->fft_n > FFT_MAX || s->fft_n < 8 ||
        num_channels < 1 || num_channels > NCH_MAX) { free(s); return NULL; }
    build_window(s);
    kf_init(&s->kf, s->fft_n / 2);
    if (!build_filterbank(s)) { free(s); return NULL; }
    s->smoothing_bits = 10;
    s->even_smoothing = (uint16_t)(0.025f * (1 << NR_BITS));
    s->odd_smoothing = (uint16_t)(0.06f * (1 << NR_BITS));
    s->min_signal_remaining = (uint16_t)(0.05f * (1 << NR_BITS));
    build_pcan(s);
    build_log_lut(s);
    s->scale_shift = 6;
    mwwo_frontend_reset(s);
    return s;
}

void mwwo_frontend_free(mwwo_frontend *s) { free(s); }

void mwwo_frontend_reset(mwwo_frontend *s) {
    s->input_used = 0;
    memset(s->input, 0, sizeof s->input);
    memset(s->win_out, 0, sizeof s->win_out);
    s->max_abs = 0;
    memset(s->fft_in, 0, sizeof s->fft_in);
    memset(s->fft_out, 0, sizeof s->fft_out);
    memset(s->work, 0, sizeof s->work);
    memset(s->estimate, 0, sizeof s->estimate);
}

/* ------------------------------------------------------------------------- */
/* per-frame stages                                                          */

static int window_process(struct mwwo_frontend *s, const int16_t *samples, size_t n, size_t *n_read) {
    const size_t WIN = (size_t)s->win, STEP = (size_t)s->step;
    size_t take = WIN - s->input_used;
    if (take > n) take = n;
    memcpy(s->input + s->input_used, samples, take * sizeof(int16_t));
    *n_read = take;
    s->input_used += take;
    if (s->input_used < WIN) return 0;

    int16_t max_abs = 0;
    for (size_t i = 0; i < WIN; ++i) {
        int16_t v = (int16_t)((((int32_t)s->input[i]) * s->coef[i]) >> WINDOW_BITS);
        s->win_out[i] = v;
        if (v < 0) v = (int16_t)(-v);  /* -(-32768) wraps back to -32768 like upstream */
        if (v > max_abs) max_abs = v;
    }
    memmove(s->input, s->input + STEP, sizeof(int16_t) * (WIN - STEP));
    s->input_used -= STEP;
    s->max_abs = max_abs;
    return 1;
}

static uint16_t sqrt32(uint32_t num) {
    if (num == 0) return 0;
    uint32_t res = 0;
    int max_bit_number = 32 - msb32(num);
    max_bit_number |= 1;
    uint32_t bit = 1U << (31 - max_bit_number);
    int iterations = (31 - max_bit_number) / 2 + 1;
    while (iterations--) {
        if (num >= res + bit) { num -= res + bit; res = (res >> 1U) + bit; }
        else res >>= 1U;
        bit >>= 2U;
    }
    if (num > res && res != 0xFFFF) ++res;
    return (uint16_t)res;
}

static uint32_t sqrt64(uint64_t num) {
    if ((num >> 32) == 0) return sqrt32((uint32_t)num);
    uint64_t res = 0;
    int max_bit_number = 64 - msb64(num);
    max_bit_number |= 1;
    uint64_t bit = 1ULL << (63 - max_bit_number);
    int iterations = (63 - max_bit_number) / 2 + 1;
    while (iterations--) {
        if (num >= res + bit) { num -= res + bit; res = (res >> 1U) + bit; }
        else res >>= 1U;
        bit >>= 2U;
    }
    if (num > res && res != 0xFFFFFFFFLL) ++res;
    return (uint32_t)res;
}

static int16_t wide_dynamic_function(uint32_t x, const int16_t *lut) {
    if (x <= 2) return lut[x];
    const int16_t interval = (int16_t)msb32(x);
    lut += 4 * interval - 6;
    const int16_t frac = (int16_t)(((interval < 11) ? (x << (11 - interval)) : (x >> (interval - 11))) & 0x3FF);
    int32_t result = ((int32_t)lut[2] * frac) >> 5;
    result += (int32_t)((uint32_t)lut[1] << 5);
    result *= frac;
    result = (result + (1 << 14)) >> 15;
    result += lut[0];
    return (int16_t)result;
}

static uint32_t pcan_shrink(uint32_t x) {
    if (x < (2u << PCAN_SNR_BITS)) return (x * x) >> (2 + 2 * PCAN_SNR_BITS - PCAN_OUT_BITS);
    return (x >> (PCAN_SNR_BITS - PCAN_OUT_BITS)) - (1u << PCAN_OUT_BITS);
}

static uint32_t log2_fraction(const uint16_t *lut, uint32_t x, uint32_t log2x) {
    int32_t frac = (int32_t)(x - (1LL << log2x));
    if (log2x < LOG_SCALE_LOG2) frac <<= LOG_SCALE_LOG2 - log2x;
    else frac >>= log2x - LOG_SCALE_LOG2;
    const uint32_t base_seg = (uint32_t)frac >> (LOG_SCALE_LOG2 - LOG_SEG_LOG2);
    const uint32_t seg_unit = (((uint32_t)1) << LOG_SCALE_LOG2) >> LOG_SEG_LOG2;
    const int32_t c0 = lut[base_seg];
    const int32_t c1 = lut[base_seg + 1];
    const int32_t seg_base = (int32_t)(seg_unit * base_seg);
    const int32_t rel_pos = ((c1 - c0) * (frac - seg_base)) >> LOG_SCALE_LOG2;
    return (uint32_t)(frac + c0 + rel_pos);
}

static uint32_t log_scaled(const uint16_t *lut, uint32_t x, uint32_t scale_shift) {
    const uint32_t integer = (uint32_t)msb32(x) - 1;
    const uint32_t fraction = log2_fraction(lut, x, integer);
    const uint32_t log2 = (integer << LOG_SCALE_LOG2) + fraction;
    const uint32_t round = (1u << LOG_SCALE_LOG2) / 2;
    const uint32_t loge = (uint32_t)((((uint64_t)LOG_COEFF) * log2 + round) >> LOG_SCALE_LOG2);
    return ((loge << scale_shift) + round) >> LOG_SCALE_LOG2;
}

int mwwo_frontend_process(mwwo_frontend *s, const int16_t *samples, size_t n, size_t *n_read, uint16_t *out) {
    size_t dummy;
    if (!n_read) n_read = &dummy;
    if (!window_process(s, samples, n, n_read)) return 0;

    /* scale so the fixed-point FFT sees as many significant bits as possible */
    const int shift = 15 - msb32((uint32_t)s->max_abs);
    s->last_shift = shift;
    int i;
    const int NCH = s->nch;
    for (i = 0; i < s->win; ++i) s->fft_in[i] = (int16_t)(uint16_t)(((uint16_t)s->win_out[i]) << shift);
    for (; i < s->fft_n; ++i) s->fft_in[i] = 0;
    kf_fftr(&s->kf, s->fft_in, s->fft_out);

    /* energy of the bins the filterbank touches */
    memset(s->last_energy, 0, sizeof s->last_energy);
    for (int b = s->start_index; b < s->end_index; ++b) {
        const int32_t re = s->fft_out[b].r, im = s->fft_out[b].i;
        s->last_energy[b] = (uint32_t)(re * re) + (uint32_t)(im * im);
    }

    /* triangular mel accumulation: running weight / unweight accumulators, 41 ranges */
    uint64_t wacc = 0, uacc = 0;
    for (int ch = 0; ch <= NCH; ++ch) {
        for (int b = s->chan_start[ch]; b < s->chan_start[ch + 1]; ++b) {
            /* upstream reads the energy through an int32_t pointer and widens it */
            const uint64_t mag = (uint64_t)(int32_t)s->last_energy[b];
            wacc += s->bin_weight[b] * mag;
            uacc += s->bin_unweight[b] * mag;
        }
        s->work[ch] = wacc;
        wacc = uacc;
        uacc = 0;
    }
    memcpy(s->last_work, s->work, sizeof s->work);

    uint32_t sig[NCH_MAX];
    for (int ch = 0; ch < NCH; ++ch) {
        sig[ch] = sqrt64(s->work[ch + 1]) >> shift;
        s->last_sqrt[ch] = sig[ch];
    }

    /* noise reduction */
    for (int ch = 0; ch < NCH; ++ch) {
        const uint32_t smoothing = ((ch & 1) == 0) ? s->even_smoothing : s->odd_smoothing;
        const uint32_t one_minus = (1u << NR_BITS) - smoothing;
        const uint32_t scaled = sig[ch] << s->smoothing_bits;
        uint32_t est = (uint32_t)((((uint64_t)scaled * smoothing) + ((uint64_t)s->estimate[ch] * one_minus)) >> NR_BITS);
        s->estimate[ch] = est;
        if (est > scaled) est = scaled;
        const uint32_t floor_v = (uint32_t)(((uint64_t)sig[ch] * s->min_signal_remaining) >> NR_BITS);
        const uint32_t sub = (scaled - est) >> s->smoothing_bits;
        sig[ch] = sub > floor_v ? sub : floor_v;
        s->last_nr[ch] = sig[ch];
    }

    /* PCAN: gain from the (just updated) noise estimate */
    for (int ch = 0; ch < NCH; ++ch) {
        const uint32_t gain = (uint32_t)wide_dynamic_function(s->estimate[ch], s->gain_lut);
        const uint32_t snr = (uint32_t)(((uint64_t)sig[ch] * gain) >> s->snr_shift);
        sig[ch] = pcan_shrink(snr);
        s->last_pcan[ch] = sig[ch];
    }

    /* log scale */
    for (int ch = 0; ch < NCH; ++ch) {
        uint32_t v = sig[ch];
        if (s->correction_bits < 0) v >>= -s->correction_bits; else v <<= s->correction_bits;
        v = (v > 1) ? log_scaled(s->log_lut, v, (uint32_t)s->scale_shift) : 0;
        out[ch] = (v < 0xFFFF) ? (uint16_t)v : 0xFFFF;
    }
    return NCH;
}

/* ------------------------------------------------------------------------- */
/* the chunk loop of generate_features_for_clip (audio_utils.py:50-64)         */

size_t mwwo_generate_features(const int16_t *audio, size_t n_samples, uint16_t *out, size_t max_rows) {
    mwwo_frontend *s = mwwo_frontend_create();   /* fresh frontend per call, audio_utils.py:52 */
    if (!s) return 0;
    size_t rows = 0;
    size_t idx = 0;                       /* byte index, as in the reference loop */
    const size_t n_bytes = n_samples * 2;
    const size_t NCH = (size_t)s->nch;
    uint16_t feat[NCH_MAX];
    while (idx + 160 * 2 < n_bytes) {     /* strict '<' : audio_utils.py:56 */
        size_t n_read = 0;
        int got = mwwo_frontend_process(s, audio + idx / 2, 160, &n_read, feat);
        idx += n_read * 2;
        if (got) {
            if (rows < max_rows) memcpy(out + rows * NCH, feat, NCH * sizeof(uint16_t));
            ++rows;
        }
    }
    mwwo_frontend_free(s);
    return rows;
}

/* stream variant: state persists in `s`; audio fed in 160-sample hops (any remainder is fed too) */
size_t mwwo_frontend_stream(mwwo_frontend *s, const int16_t *audio, size_t n_samples, uint16_t *out, size_t max_rows) {
    size_t rows = 0, pos = 0;
    const size_t NCH = (size_t)s->nch, HOP = (size_t)s->step;
    uint16_t feat[NCH_MAX];
    while (pos < n_samples) {
        size_t n = n_samples - pos; if (n > HOP) n = HOP;
        size_t n_read = 0;
        int got = mwwo_frontend_process(s, audio + pos, n, &n_read, feat);
        pos += n_read;
        if (got) {
            if (rows < max_rows) memcpy(out + rows * NCH, feat, NCH * sizeof(uint16_t));
            ++rows;
        }
    }
    return rows;
}

/* ------------------------------------------------------------------------- */
/* table / tap accessors for the known-answer tests                           */

void mwwo_frontend_tables(const mwwo_frontend *s, int16_t *window480, int16_t *bin_channel257,
                          int16_t *bin_weight257, int16_t *bin_unweight257, int16_t *chan_start42,
                          int16_t *gain_lut125, uint16_t *log_lut129, int16_t *twiddles512,
                          int16_t *super256, int32_t *scalars8) {
    if (window480) memcpy(window480, s->coef, sizeof s->coef);
    if (bin_channel257) memcpy(bin_channel257, s->bin_channel, sizeof s->bin_channel);
    if (bin_weight257) memcpy(bin_weight257, s->bin_weight, sizeof s->bin_weight);
    if (bin_unweight257) memcpy(bin_unweight257, s->bin_unweight, sizeof s->bin_unweight);
    if (chan_start42) memcpy(chan_start42, s->chan_start, sizeof s->chan_start);
    if (gain_lut125) memcpy(gain_lut125, s->gain_lut, sizeof s->gain_lut);
    if (log_lut129) memcpy(log_lut129, s->log_lut, 129 * sizeof(uint16_t));
    if (twiddles512) memcpy(twiddles512, s->kf.tw, sizeof s->kf.tw);
    if (super256) memcpy(super256, s->kf.super, sizeof s->kf.super);
    if (scalars8) {
        scalars8[0] = s->start_index; scalars8[1] = s->end_index;
        scalars8[2] = s->even_smoothing; scalars8[3] = s->odd_smoothing;
        scalars8[4] = s->min_signal_remaining; scalars8[5] = s->snr_shift;
        scalars8[6] = s->correction_bits; scalars8[7] = s->scale_shift;
    }
}

void mwwo_frontend_taps(const mwwo_frontend *s, int32_t *shift, int16_t *fft_in512, int16_t *fft_out514,
                        uint32_t *energy257, uint64_t *work41, uint32_t *sqrt40, uint32_t *nr40,
                        uint32_t *pcan40, uint32_t *estimate40) {
    if (shift) *shift = s->last_shift;
    if (fft_in512) memcpy(fft_in512, s->fft_in, sizeof s->fft_in);
    if (fft_out514) memcpy(fft_out514, s->fft_out, sizeof s->fft_out);
    if (energy257) memcpy(energy257, s->last_energy, sizeof s->last_energy);
    if (work41) memcpy(work41, s->last_work, sizeof s->last_work);
    if (sqrt40) memcpy(sqrt40, s->last_sqrt, sizeof s->last_sqrt);
    if (nr40) memcpy(nr40, s->last_nr, sizeof s->last_nr);
    if (pcan40) memcpy(pcan40, s->last_pcan, sizeof s->last_pcan);
    if (estimate40) memcpy(estimate40, s->estimate, sizeof s->estimate);
}

/* runtime sizes: {sample_rate, window, step, fft_size, num_channels, start_index, end_index, 0} */
void mwwo_frontend_config(const mwwo_frontend *s, int32_t *cfg8) {
    cfg8[0] = s->sample_rate; cfg8[1] = s->win; cfg8[2] = s->step; cfg8[3] = s->fft_n;
    cfg8[4] = s->nch; cfg8[5] = s->start_index; cfg8[6] = s->end_index; cfg8[7] = 0;
}
/* windowed frame of the last produced row (upstream WindowState.output) and its max |value| */
void mwwo_frontend_window_tap(const mwwo_frontend *s, int16_t *win_out512, int32_t *max_abs) {
    if (win_out512) memcpy(win_out512, s->win_out, sizeof s->win_out);
    if (max_abs) *max_abs = s->max_abs;
}

void mwwo_frontend_get_state(const mwwo_frontend *s, int16_t *input480, int32_t *input_used, uint32_t *estimate40) {
    if (input480) memcpy(input480, s->input, sizeof s->input);
    if (input_used) *input_used = (int32_t)s->input_used;
    if (estimate40) memcpy(estimate40, s->estimate, sizeof s->estimate);
}

/* raw entry points so tests can hit single stages with adversarial operands */
uint32_t mwwo_sqrt64(uint64_t x) { return sqrt64(x); }
int32_t mwwo_wdf(const mwwo_frontend *s, uint32_t x) { return wide_dynamic_function(x, s->gain_lut); }
uint32_t mwwo_pcan_shrink(uint32_t x) { return pcan_shrink(x); }
uint32_t mwwo_log_scaled(const mwwo_frontend *s, uint32_t x) { return log_scaled(s->log_lut, x, (uint32_t)s->scale_shift); }
void mwwo_fftr(const mwwo_frontend *s, const int16_t *in512, int16_t *out514) { kf_fftr(&s->kf, in512, (cpx *)out514); }

// mww_tables.h -- host-built constant tables of the fixed-point micro-frontend, in the packed
// layouts the sm_100a kernels consume.  Replaces the table construction that
// pymicro_features.MicroFrontend() performs when the reference instantiates it at
// microwakeword/audio/audio_utils.py:52 (window / KissFFT twiddles / mel filterbank weights /
// PCAN gain LUT / log LUT; algorithm per SURVEY.md Appendix B).
#pragma once

#include <stdint.h>

#include <vector>

#include "mww_common.h"

namespace mww {

constexpr int kFbLanes = 16;       // lanes that cooperate on one frame
constexpr int kFbSlots = 4;        // max channels one lane accumulates
constexpr int kFbCoefMax = 1024;   // capacity of the shared-memory copy of the span coefficients
// Trip counts of the balanced filterbank schedule (mww_tables.cc LPT assignment), padded to even so the
// Q12 coefficients can be fetched two per 32-bit load.  mww_create verifies the table builder agrees.
constexpr int kFbLen[kFbSlots] = {28, 12, 6, 0};
MWW_HD constexpr int fb_len(int s) { return s == 0 ? 28 : (s == 1 ? 12 : (s == 2 ? 6 : 0)); }
static_assert(fb_len(0) == kFbLen[0] && fb_len(1) == kFbLen[1] && fb_len(2) == kFbLen[2] && fb_len(3) == kFbLen[3], "fb_len");

struct FbSlot {
    int16_t ch;        // output channel 0..39, -1 = unused slot
    int16_t bin0;      // first FFT bin of the span
    int16_t n;         // bins in the span
    int16_t coef_off;  // offset into fb_coef (span is zero padded to the slot's uniform length)
};

// Tables passed BY VALUE as a kernel parameter (lives in the constant bank; uniform reads are free).
struct FrontendParams {
    const uint32_t *win_pairs;   // [240] window coefficients, two Q12 int16 per word
    const uint32_t *tw;          // [256] exp(-2*pi*i*k/256) Q15, (re | im << 16)
    const uint32_t *super_tw;    // [128] real-FFT post-pass twiddles, same packing
    const int16_t *fb_coef;      // filterbank span coefficients (Q12), zero padded
    const FbSlot *fb_slots;      // [16][kFbSlots]
    const int16_t *gain_lut;     // [128] PCAN wide-dynamic-function LUT (125 used)
    const uint16_t *log_lut;     // [132] log2 correction LUT (129 used)
    int16_t tw2[9][2];           // stage-2 twiddles tw[16k], tw[32k], tw[48k] for k = 1..3
    int32_t fb_slot_len[kFbSlots];
};

struct HostTables {
    // packed, kernel-facing
    uint32_t win_pairs[240];
    uint32_t tw[256];
    uint32_t super_tw[128];
    std::vector<int16_t> fb_coef;
    FbSlot fb_slots[kFbLanes][kFbSlots];
    int32_t fb_slot_len[kFbSlots];
    int16_t gain_lut[128];
    uint16_t log_lut[132];
    int16_t tw2[9][2];
    // logical, test-facing
    int16_t window[480];
    int16_t bin_channel[257], bin_weight[257], bin_unweight[257];
    int16_t chan_start[42];
    int start_index, end_index;
    bool ok;
};

// noise-reduction / PCAN / log scalars
constexpr int kSmoothingBits = 10;
constexpr uint32_t kEvenSmoothing = 409;       // int(0.025 * 2^14)
constexpr uint32_t kOddSmoothing = 983;        // int(0.06  * 2^14)
constexpr uint32_t kMinSignalRemaining = 819;  // int(0.05  * 2^14)
constexpr int kNoiseBits = 14;
constexpr int kPcanSnrShift = 6;               // gain_bits 21 - correction 3 - snr bits 12
constexpr int kLogCorrectionBits = 3;          // msb(512) - 1 - 12/2
constexpr int kLogScaleShift = 6;

void build_host_tables(HostTables *t);

}  // namespace mww

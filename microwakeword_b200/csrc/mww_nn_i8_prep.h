// mww_nn_i8_prep.h -- host-side preparation of the int8 tensor-core operands (see mww_nn_i8_dev.cuh):
// K-contiguous (transposed, padded) weight rows and zero-point-folded biases.  Shared by mww_create and by
// the host emulation in tests/host_emul so both run on identical operands.
#pragma once

#include <stdint.h>

#include <vector>

#include "mww_nn_i8_dev.cuh"

namespace mww {

struct I8MmaOperands {
    std::vector<int8_t> w0t;          // [32][kW0Pitch]
    std::vector<int8_t> pwt[4];       // [64][kPwPitch]
    std::vector<int32_t> b0f;         // [32]
    std::vector<int32_t> pw_bf[4];    // [64]
    std::vector<int8_t> qlut;         // [65536] (build_feature_qlut; the host emulation keeps it here)
};

// NnWeightsI8::qlut: the input quantisation of every possible uint16 feature value, by the kernels' own expression
inline void build_feature_qlut(float in_scale, int32_t zp_in, std::vector<int8_t> *out) {
    out->resize(65536);
    for (int u = 0; u < 65536; ++u) (*out)[u] = (int8_t)nnq_quantize((float)(uint16_t)u * kFeatureScale, in_scale, zp_in);
}

// w0: [200][32] int8, b0: [32]; pw_w[L]: [cin][64], pw_b[L]: [64]; zp: the 12 activation zero points
inline void build_i8_mma_operands(const int8_t *w0, const int32_t *b0, const int8_t *const pw_w[4], const int32_t *const pw_b[4],
                                  const int32_t *zp, I8MmaOperands *out) {
    out->w0t.assign((size_t)32 * kW0Pitch, 0);
    out->b0f.assign(32, 0);
    for (int n = 0; n < 32; ++n) {
        int32_t colsum = 0;
        for (int k = 0; k < 200; ++k) {
            const int8_t w = w0[k * 32 + n];
            out->w0t[(size_t)n * kW0Pitch + k] = w;
            colsum += w;
        }
        out->b0f[n] = b0[n] - zp[0] * colsum;
    }
    static const int cin[4] = {32, 64, 64, 64};
    for (int L = 0; L < 4; ++L) {
        out->pwt[L].assign((size_t)64 * kPwPitch, 0);
        out->pw_bf[L].assign(64, 0);
        for (int n = 0; n < 64; ++n) {
            int32_t colsum = 0;
            for (int k = 0; k < cin[L]; ++k) {
                const int8_t w = pw_w[L][k * 64 + n];
                out->pwt[L][(size_t)n * kPwPitch + k] = w;
                colsum += w;
            }
            out->pw_bf[L][n] = pw_b[L][n] - zp[2 + 2 * L] * colsum;
        }
    }
}

}  // namespace mww

// mww_nn_i8_dev.cuh -- integer (int8-quantised) streaming MixedNet, clip formulation, per-thread phases.
//
// Executes the graph a quantised reference model runs inside tf.lite.Interpreter.invoke()
// (microwakeword/inference.py:110-121) under the conversion contract of
// microwakeword/utils.py:289-348: int8 activations (per-tensor, asymmetric), int8 weights
// (per-output-channel symmetric; per-tensor for the dense head), int32 biases, int8 ring-buffer
// state variables (:333), int8 input at scale 26/255 (:308-313), uint8 output (:338), requantised
// with TFLite's SaturatingRoundingDoublingHighMul + RoundingDivideByPOT.
// Result is bit-exact with oracle/mixednet.c::mwwo_mixednet_step_int8.
//
// Same time-parallel shared-memory geometry as the fp32 path (mww_nn_dev.cuh); activations are
// kept as int32 words holding (q - zero_point), so a zero word is a real 0 and ring padding,
// depthwise taps and pointwise inputs need no per-use zero-point subtraction.
#pragma once

#include "mww_common.h"
#include "mww_nn_dev.cuh"

namespace mww {

struct NnWeightsI8 {
    const int8_t *w0; const int32_t *b0, *m0, *s0;              // first conv [5][40][32]
    const int8_t *dw_w[4]; const int32_t *dw_b[4], *dw_m[4], *dw_s[4];
    const int8_t *pw_w[4]; const int32_t *pw_b[4], *pw_m[4], *pw_s[4];
    const int8_t *head_w; const int8_t *lut;
    int32_t head_bias, head_mult, head_shift;
    int32_t zp[12];          // [in, c0, d1, p1, d2, p2, d3, p3, d4, p4, fc, prob]
    float in_scale;
};

// zero point of the tensor buffered by ring L (0..3: block inputs, 4: head input)
MWW_HD int32_t nnq_ring_zp(const NnWeightsI8 &W, int L) { return W.zp[1 + 2 * L]; }

MWW_HD int32_t srdhm(int32_t a, int32_t b) {
    if (a == INT32_MIN && b == INT32_MIN) return INT32_MAX;
    const int64_t ab = (int64_t)a * (int64_t)b;
    const int64_t t = ab + (ab >= 0 ? (1 << 30) : (1 - (1 << 30)));
    // C++ division by 2^31 truncates toward zero
    return (int32_t)(t >= 0 ? (t >> 31) : -((-t) >> 31));
}
MWW_HD int32_t rounding_divide_by_pot(int32_t x, int exponent) {
    const int32_t mask = (int32_t)((1ll << exponent) - 1);
    const int32_t rem = x & mask;
    const int32_t thr = (mask >> 1) + (x < 0 ? 1 : 0);
    return (x >> exponent) + (rem > thr ? 1 : 0);
}
MWW_HD int32_t mbqm(int32_t x, int32_t mult, int32_t shift) {
    const int left = shift > 0 ? shift : 0, right = shift > 0 ? 0 : -shift;
    return rounding_divide_by_pot(srdhm((int32_t)((uint32_t)x << left), mult), right);
}
// requantise an accumulator; returns (q_out - zp_out) so it can be stored pre-subtracted
MWW_HD int32_t requant_rel(int32_t acc, int32_t mult, int32_t shift, int32_t zp_out, bool relu) {
    int32_t y = mbqm(acc, mult, shift) + zp_out;
    const int32_t lo = relu ? zp_out : -128;
    y = y < lo ? lo : y;
    y = y > 127 ? 127 : y;
    return y - zp_out;
}

struct NnInputI8 {
    const int8_t *ring0;      // [2][40] raw int8
    const int8_t *pend;       // [2][40]
    int n_pend;
    const void *rows;
    int n_rows;
    int row_type;             // 0 uint16, 1 float32, 2 int8
};

// Model.quantize_input_data (inference.py:127-147): float divide, add zero point, astype(int8)
MWW_HD int32_t nnq_quantize(float x, float scale, int32_t zp) { return (int32_t)(int8_t)(int32_t)(x / scale + (float)zp); }

// virtual feature row element as (q - zp_in)
MWW_HD int32_t nnq_virtual_row(const NnInputI8 &in, const NnWeightsI8 &W, int vr, int f) {
    int32_t q;
    if (vr < 0) q = in.ring0[(2 + vr) * kNumChannels + f];
    else if (vr < in.n_pend) q = in.pend[vr * kNumChannels + f];
    else {
        const long long e = (long long)(vr - in.n_pend) * kNumChannels + f;
        if (in.row_type == 2) q = static_cast<const int8_t *>(in.rows)[e];
        else {
            const float x = in.row_type == 1 ? static_cast<const float *>(in.rows)[e]
                                             : (float)static_cast<const uint16_t *>(in.rows)[e] * kFeatureScale;
            q = nnq_quantize(x, W.in_scale, W.zp[0]);
        }
    }
    return q - W.zp[0];
}

template <int L>
MWW_HD void nnq_load_state_l(int tid, int32_t *sm, const int8_t *state, const NnWeightsI8 &W) {
    constexpr NnLayerGeom g = kGeom[L];
    const int8_t *src = state + kStateOff[L + 1];
    const int32_t zp = nnq_ring_zp(W, L);
    for (int e = tid; e < g.ring * g.cin; e += kNnThreads) {
        const int r = e / g.cin, c = e - r * g.cin;
        sm[g.off + c * g.ld + (g.hp - g.ring) + r] = (int32_t)src[e] - zp;
    }
}
MWW_HD void nnq_load_state(int tid, int32_t *sm, const int8_t *state, const NnWeightsI8 &W) {
    nnq_load_state_l<0>(tid, sm, state, W); nnq_load_state_l<1>(tid, sm, state, W); nnq_load_state_l<2>(tid, sm, state, W);
    nnq_load_state_l<3>(tid, sm, state, W); nnq_load_state_l<4>(tid, sm, state, W);
}

MWW_HD void nnq_load_features(int tid, int32_t *sm, const NnInputI8 &in, const NnWeightsI8 &W, int step0, int n) {
    int32_t *feat = sm + kXFloats + kDFloats;
    for (int e = tid; e < 5 * kNumChannels * kTT; e += kNnThreads) {
        const int t = e % kTT;
        const int f = (e / kTT) % kNumChannels;
        const int j = e / (kTT * kNumChannels);
        int32_t v = 0;
        if (t < n) v = nnq_virtual_row(in, W, 3 * (step0 + t) + j - 2, f);
        feat[(j * kNumChannels + f) * kTT + t] = v;
    }
}

MWW_HD void nnq_first_conv(int tid, int32_t *sm, const NnWeightsI8 &W) {
    const int32_t *feat = sm + kXFloats + kDFloats;
    const int o = tid & 31, q = tid >> 5;
    int32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0;
    for (int k = 0; k < 5 * kNumChannels; ++k) {
        const int32_t w = W.w0[k * 32 + o];
        const int32_t *x = feat + k * kTT + 8 * q;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += w * x[i];
    }
    const int32_t b = W.b0[o], m = W.m0[o], s = W.s0[o];
    int32_t *dst = sm + kGeom[0].off + o * kGeom[0].ld + kGeom[0].hp + 8 * q;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = requant_rel(acc[i] + b, m, s, W.zp[1], true);
}

template <int L>
MWW_HD void nnq_depthwise(int tid, int32_t *sm, const NnWeightsI8 &W) {
    constexpr NnLayerGeom g = kGeom[L];
    constexpr int TPC = kNnThreads / g.cin;
    constexpr int TN = kTT / TPC;
    const int c = tid % g.cin, part = tid / g.cin;
    const int t0 = part * TN;
    int32_t w[g.kmax];
#pragma unroll
    for (int j = 0; j < g.kmax; ++j) w[j] = W.dw_w[L][j * g.cin + c];
    int32_t acc[TN];
#pragma unroll
    for (int i = 0; i < TN; ++i) acc[i] = 0;
    const int32_t *x = sm + g.off + c * g.ld + (g.hp - (g.kmax - 1)) + t0;
#pragma unroll
    for (int i = 0; i < TN + g.kmax - 1; ++i) {
        const int32_t xv = x[i];
#pragma unroll
        for (int tt = 0; tt < TN; ++tt) {
            const int j = i - tt;
            if (j >= 0 && j < g.kmax) acc[tt] += w[j] * xv;
        }
    }
    const int32_t b = W.dw_b[L][c], m = W.dw_m[L][c], s = W.dw_s[L][c];
    int32_t *d = sm + kXFloats + c * kDLd + t0;
#pragma unroll
    for (int i = 0; i < TN; ++i) d[i] = requant_rel(acc[i] + b, m, s, W.zp[2 + 2 * L], false);
}

template <int L>
MWW_HD void nnq_pointwise(int tid, int32_t *sm, const NnWeightsI8 &W) {
    constexpr int cin = kGeom[L].cin;
    constexpr NnLayerGeom gn = kGeom[L + 1];
    const int o = tid & 63, h = tid >> 6;
    int32_t acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    const int32_t *d = sm + kXFloats + 16 * h;
    for (int k = 0; k < cin; ++k) {
        const int32_t w = W.pw_w[L][k * 64 + o];
        const int32_t *x = d + k * kDLd;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += w * x[i];
    }
    const int32_t b = W.pw_b[L][o], m = W.pw_m[L][o], s = W.pw_s[L][o];
    int32_t *dst = sm + gn.off + o * gn.ld + gn.hp + 16 * h;
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i] = requant_rel(acc[i] + b, m, s, W.zp[3 + 2 * L], true);
}

MWW_HD void nnq_head_partial(int tid, int32_t *sm, const NnWeightsI8 &W) {
    constexpr NnLayerGeom g = kGeom[4];
    const int c = tid & 63, h = tid >> 6;
    int32_t w[17];
#pragma unroll
    for (int j = 0; j < 17; ++j) w[j] = W.head_w[j * 64 + c];
    int32_t acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    const int32_t *x = sm + g.off + c * g.ld + (g.hp - 16) + 16 * h;
#pragma unroll
    for (int i = 0; i < 16 + 16; ++i) {
        const int32_t xv = x[i];
#pragma unroll
        for (int tt = 0; tt < 16; ++tt) {
            const int j = i - tt;
            if (j >= 0 && j < 17) acc[tt] += w[j] * xv;
        }
    }
    int32_t *d = sm + kXFloats + c * kDLd + 16 * h;
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = acc[i];
}

// FULLY_CONNECTED requant -> LOGISTIC LUT -> QUANTIZE to uint8 -> Model.dequantize_output_data (/255)
MWW_HD void nnq_head_finish(int tid, int32_t *sm, const NnWeightsI8 &W, int n, float *probs_out) {
    if (tid >= kTT || tid >= n) return;
    const int32_t *d = sm + kXFloats + tid;
    int32_t acc = 0;
    for (int c = 0; c < 64; ++c) acc += d[c * kDLd];
    const int32_t logit = requant_rel(acc + W.head_bias, W.head_mult, W.head_shift, W.zp[10], false) + W.zp[10];
    const int32_t out_u8 = (int32_t)W.lut[(uint8_t)(int8_t)logit] + 128;
    probs_out[tid] = (1.0f / 255.0f) * (float)out_u8;                 // inference.py:162-170
}

template <int L>
MWW_HD void nnq_store_state_l(int tid, const int32_t *sm, int8_t *state, const NnWeightsI8 &W) {
    constexpr NnLayerGeom g = kGeom[L];
    int8_t *dst = state + kStateOff[L + 1];
    const int32_t zp = nnq_ring_zp(W, L);
    for (int e = tid; e < g.ring * g.cin; e += kNnThreads) {
        const int r = e / g.cin, c = e - r * g.cin;
        dst[e] = (int8_t)(sm[g.off + c * g.ld + (g.hp - g.ring) + r] + zp);
    }
}

struct NnTailI8 { int8_t ring_new[2], pend_new[2]; };
MWW_HD void nnq_tail_read(int tid, const NnInputI8 &in, const NnWeightsI8 &W, int n_steps, int n_virtual_rows, NnTailI8 &t) {
    const int consumed = 3 * n_steps;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = tid + q * kNnThreads;
        t.ring_new[q] = 0; t.pend_new[q] = 0;
        if (e < 2 * kNumChannels) {
            const int r = e / kNumChannels, f = e - r * kNumChannels;
            t.ring_new[q] = (int8_t)(nnq_virtual_row(in, W, consumed - 2 + r, f) + W.zp[0]);
            const int vr = consumed + r;
            t.pend_new[q] = vr < n_virtual_rows ? (int8_t)(nnq_virtual_row(in, W, vr, f) + W.zp[0]) : (int8_t)W.zp[0];
        }
    }
}
MWW_HD void nnq_tail_write(int tid, const int32_t *sm, int8_t *state, int8_t *pend_out, const NnWeightsI8 &W, const NnTailI8 &t) {
    nnq_store_state_l<0>(tid, sm, state, W); nnq_store_state_l<1>(tid, sm, state, W); nnq_store_state_l<2>(tid, sm, state, W);
    nnq_store_state_l<3>(tid, sm, state, W); nnq_store_state_l<4>(tid, sm, state, W);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = tid + q * kNnThreads;
        if (e < 2 * kNumChannels) { state[e] = t.ring_new[q]; pend_out[e] = t.pend_new[q]; }
    }
}

// value every state byte takes after a reset: the zero point of the tensor it buffers
MWW_HD int8_t nnq_reset_value(const NnWeightsI8 &W, int e /* 0..4175 */) {
    // segment ends: first-conv ring 80, block rings 208 / 848 / 1744 / 3152, head ring 4176
    if (e < 80) return (int8_t)W.zp[0];
    if (e < 208) return (int8_t)W.zp[1];
    if (e < 848) return (int8_t)W.zp[3];
    if (e < 1744) return (int8_t)W.zp[5];
    if (e < 3152) return (int8_t)W.zp[7];
    return (int8_t)W.zp[9];
}

}  // namespace mww

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

// de-interleave the chunk's rows into three planes of (q - zp_in) words (same indexing as the fp32 path)
MWW_HD void nnq_load_features(int tid, int32_t *sm, const NnInputI8 &in, const NnWeightsI8 &W, int step0, int n) {
    int32_t *feat = sm + kXFloats + kDFloats;
    const int n_q = 3 * n + 2;
    for (int e = tid; e < n_q * kNumChannels; e += kNnThreads) {
        const int q = e / kNumChannels, f = e - q * kNumChannels;
        feat[((q % 3) * kNumChannels + f) * kUS + q / 3] = nnq_virtual_row(in, W, 3 * step0 + q - 2, f);
    }
}

// first conv, K split over two thread groups (taps 0..2 | 3..4); integer partial sums meet in D
MWW_HD void nnq_first_conv_a(int tid, int32_t *sm, const NnWeightsI8 &W, int32_t (&acc)[2][4]) {
    const int32_t *feat = sm + kXFloats + kDFloats;
    const int half = tid >= 144, r = tid - 144 * half;
    const int o0 = 2 * (r & 15), t0 = 4 * (r >> 4);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0;
    const int j_begin = half ? 3 : 0, j_end = half ? 5 : 3;
    for (int j = j_begin; j < j_end; ++j) {
        const int32_t *plane = feat + (j % 3) * kNumChannels * kUS + t0 + j / 3;
        const int8_t *w = W.w0 + j * kNumChannels * 32 + o0;
#pragma unroll 8
        for (int f = 0; f < kNumChannels; ++f) {
            const int32_t w0 = w[f * 32], w1 = w[f * 32 + 1];
            const int32_t *x = plane + f * kUS;
#pragma unroll
            for (int q = 0; q < 4; ++q) { acc[0][q] += w0 * x[q]; acc[1][q] += w1 * x[q]; }
        }
    }
    if (half) {
        int32_t *part = sm + kXFloats;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) part[(o0 + i) * kDLd + t0 + q] = acc[i][q];
    }
}
MWW_HD void nnq_first_conv_b(int tid, int32_t *sm, const NnWeightsI8 &W, const int32_t (&acc)[2][4]) {
    if (tid >= 144) return;
    const int o0 = 2 * (tid & 15), t0 = 4 * (tid >> 4);
    const int32_t *part = sm + kXFloats;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int32_t b = W.b0[o0 + i], m = W.m0[o0 + i], s = W.s0[o0 + i];
        int32_t *dst = sm + kGeom[0].off + (o0 + i) * kGeom[0].ld + kGeom[0].hp + t0;
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = requant_rel(acc[i][q] + part[(o0 + i) * kDLd + t0 + q] + b, m, s, W.zp[1], true);
    }
}

template <int L>
MWW_HD void nnq_stage_pw_weights(int tid, int32_t *sm, const NnWeightsI8 &W) {
    int32_t *wsm = sm + kXFloats + kDFloats;
    constexpr int n = kGeom[L].cin * 64;
    for (int e = tid; e < n; e += kNnThreads) wsm[e] = W.pw_w[L][e];
}

template <int L, int K, int TN>
MWW_HD void nnq_depthwise_k(int c, int t0, int32_t *sm, const NnWeightsI8 &W) {
    constexpr NnLayerGeom g = kGeom[L];
    int32_t w[K];
#pragma unroll
    for (int j = 0; j < K; ++j) w[j] = W.dw_w[L][(g.kmax - K + j) * g.cin + c];
    int32_t acc[TN];
#pragma unroll
    for (int i = 0; i < TN; ++i) acc[i] = 0;
    const int32_t *x = sm + g.off + c * g.ld + (g.hp - (K - 1)) + t0;
#pragma unroll
    for (int i = 0; i < TN + K - 1; ++i) {
        const int32_t xv = x[i];
#pragma unroll
        for (int tt = 0; tt < TN; ++tt) {
            const int j = i - tt;
            if (j >= 0 && j < K) acc[tt] += w[j] * xv;
        }
    }
    const int32_t b = W.dw_b[L][c], m = W.dw_m[L][c], s = W.dw_s[L][c];
    int32_t *d = sm + kXFloats + c * kDLd + t0;
#pragma unroll
    for (int i = 0; i < TN; ++i) d[i] = requant_rel(acc[i] + b, m, s, W.zp[2 + 2 * L], false);
}

template <int L>
MWW_HD void nnq_depthwise(int tid, int32_t *sm, const NnWeightsI8 &W) {
    constexpr NnLayerGeom g = kGeom[L];
    if (L == 0) {
        nnq_depthwise_k<L, g.kmax, 4>(tid & 31, 4 * (tid >> 5), sm, W);
    } else {
        if (tid >= 256) return;
        const int c = tid & 63, t0 = 9 * (tid >> 6);
        if (L == 1) { if (c < 32) nnq_depthwise_k<L, 7, 9>(c, t0, sm, W); else nnq_depthwise_k<L, 11, 9>(c, t0, sm, W); }
        else if (L == 2) { if (c < 32) nnq_depthwise_k<L, 9, 9>(c, t0, sm, W); else nnq_depthwise_k<L, 15, 9>(c, t0, sm, W); }
        else nnq_depthwise_k<L, g.kmax, 9>(c, t0, sm, W);
    }
}

template <int L>
MWW_HD void nnq_pointwise(int tid, int32_t *sm, const NnWeightsI8 &W) {
    constexpr int cin = kGeom[L].cin;
    constexpr NnLayerGeom gn = kGeom[L + 1];
    const int o0 = 4 * (tid & 15), t0 = 2 * (tid >> 4);
    const int32_t *wsm = sm + kXFloats + kDFloats + o0;
    const int32_t *d = sm + kXFloats + t0;
    int32_t acc[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[i][0] = 0; acc[i][1] = 0; }
#pragma unroll 8
    for (int k = 0; k < cin; ++k) {
        const int32_t *w = wsm + k * 64;
        const int32_t *x = d + k * kDLd;
        const int32_t x0 = x[0], x1 = x[1];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] += w[i] * x0; acc[i][1] += w[i] * x1; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int32_t b = W.pw_b[L][o0 + i], m = W.pw_m[L][o0 + i], s = W.pw_s[L][o0 + i];
        int32_t *dst = sm + gn.off + (o0 + i) * gn.ld + gn.hp + t0;
        dst[0] = requant_rel(acc[i][0] + b, m, s, W.zp[3 + 2 * L], true);
        dst[1] = requant_rel(acc[i][1] + b, m, s, W.zp[3 + 2 * L], true);
    }
}

MWW_HD void nnq_head_partial(int tid, int32_t *sm, const NnWeightsI8 &W) {
    constexpr NnLayerGeom g = kGeom[4];
    if (tid >= 256) return;
    const int c = tid & 63, t0 = 9 * (tid >> 6);
    int32_t w[17];
#pragma unroll
    for (int j = 0; j < 17; ++j) w[j] = W.head_w[j * 64 + c];
    int32_t acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0;
    const int32_t *x = sm + g.off + c * g.ld + (g.hp - 16) + t0;
#pragma unroll
    for (int i = 0; i < 9 + 16; ++i) {
        const int32_t xv = x[i];
#pragma unroll
        for (int tt = 0; tt < 9; ++tt) {
            const int j = i - tt;
            if (j >= 0 && j < 17) acc[tt] += w[j] * xv;
        }
    }
    int32_t *d = sm + kXFloats + c * kDLd + t0;
#pragma unroll
    for (int i = 0; i < 9; ++i) d[i] = acc[i];
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

struct NnTailI8 { int8_t ring_new, pend_new; };
MWW_HD void nnq_tail_read(int tid, const NnInputI8 &in, const NnWeightsI8 &W, int n_steps, int n_virtual_rows, NnTailI8 &t) {
    const int consumed = 3 * n_steps;
    t.ring_new = 0; t.pend_new = 0;
    if (tid < 2 * kNumChannels) {
        const int r = tid / kNumChannels, f = tid - r * kNumChannels;
        t.ring_new = (int8_t)(nnq_virtual_row(in, W, consumed - 2 + r, f) + W.zp[0]);
        const int vr = consumed + r;
        t.pend_new = vr < n_virtual_rows ? (int8_t)(nnq_virtual_row(in, W, vr, f) + W.zp[0]) : (int8_t)W.zp[0];
    }
}
MWW_HD void nnq_tail_write(int tid, const int32_t *sm, int8_t *state, int8_t *pend_out, const NnWeightsI8 &W, const NnTailI8 &t) {
    nnq_store_state_l<0>(tid, sm, state, W); nnq_store_state_l<1>(tid, sm, state, W); nnq_store_state_l<2>(tid, sm, state, W);
    nnq_store_state_l<3>(tid, sm, state, W); nnq_store_state_l<4>(tid, sm, state, W);
    if (tid < 2 * kNumChannels) { state[tid] = t.ring_new; pend_out[tid] = t.pend_new; }
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

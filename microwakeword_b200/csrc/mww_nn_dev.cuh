// mww_nn_dev.cuh -- streaming MixedNet (fp32), time-parallel "clip" formulation, per-thread phases.
//
// Replaces what tf.lite.Interpreter.invoke() executes per 30 ms step at
// microwakeword/inference.py:113-119, i.e. the streaming graph of microwakeword/mixednet.py:307-386
// (first strided conv, four MixConv blocks = ring-buffered grouped depthwise conv + 1x1 pointwise
// with folded BatchNorm + ReLU, 17-row ring, Flatten, Dense(1), sigmoid) with the ring semantics of
// microwakeword/layers/stream.py:581-595.
//
// B200-first design: every conv in the graph is a causal 'valid' FIR along time, so a whole chunk of
// TT model steps of one stream is evaluated layer by layer as small dense problems whose M dimension
// is TIME (README.md:27-28 equivalence; the ring state supplies the k-1 rows of history).  The ring
// buffers live in shared memory for the entire clip and touch HBM once per call, instead of the
// reference's 4 176-element state shuffle on every step (stream.py:586-590).
//
// Shared-memory activation layout is channel-major [c][LD] with time contiguous; columns
// [HP-R, HP) hold the ring history, columns [HP, HP+TT) the chunk's new outputs (HP = R rounded up to
// a multiple of 4 so float4 accesses stay aligned; LD/4 is odd to spread 128-bit accesses over banks).
#pragma once

#include "mww_common.h"

#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

namespace mww {

constexpr int kNnThreads = 128;
constexpr int kTT = 32;            // model steps per chunk

// okay_nabu MixedNet (notebooks/basic_training_notebook.ipynb:503-509)
struct ArchOkayNabu {
    static constexpr int C0 = 32, K0 = 5, STRIDE = 3, NB = 4, CP = 64, HEAD = 17;
    static constexpr int ring0_rows = 2;
};

struct NnLayerGeom { int cin, kmax, ring, hp, ld, off; };   // off: float offset of the buffer in smem

// geometry of the five ring-carrying activation buffers: inputs of block 0..3 and of the head
constexpr NnLayerGeom kGeom[5] = {
    {32, 5, 4, 4, 36, 0},
    {64, 11, 10, 12, 44, 32 * 36},
    {64, 15, 14, 16, 52, 32 * 36 + 64 * 44},
    {64, 23, 22, 24, 60, 32 * 36 + 64 * 44 + 64 * 52},
    {64, 17, 16, 16, 52, 32 * 36 + 64 * 44 + 64 * 52 + 64 * 60},
};
constexpr int kXFloats = 32 * 36 + 64 * 44 + 64 * 52 + 64 * 60 + 64 * 52;   // 14464
constexpr int kDLd = 36;
constexpr int kDFloats = 64 * kDLd;
constexpr int kFeatFloats = 5 * kNumChannels * kTT;   // im2col'ed feature planes [j][f][t]
constexpr int kNnSmemFloats = kXFloats + kDFloats + kFeatFloats;
constexpr int kNnSmemBytes = kNnSmemFloats * 4;       // 92.4 KB -> 2 CTAs / SM

// per-stream state layout in HBM (floats), oldest row first, [row][channel] -- identical to the oracle
constexpr int kStateOff[6] = {0, 80, 80 + 128, 80 + 128 + 640, 80 + 128 + 640 + 896, 80 + 128 + 640 + 896 + 1408};
constexpr int kStateFloats = 4176;

struct NnWeightsF32 {
    const float *w0;          // [5][40][32]
    const float *dw_w[4];     // [kmax][cin] zero padded at the front for the shorter MixConv kernels
    const float *dw_b[4];     // [cin]
    const float *pw_w[4];     // [cin][64]   BatchNorm folded
    const float *pw_b[4];     // [64]
    const float *head_w;      // [17][64]
    const float *head_b;      // [1]
};

// input feature rows: a virtual sequence  pend[0..n_pend) ++ rows[0..n_rows)
struct NnInput {
    const float *ring0;       // [2][40] first-conv ring (rows -2, -1)
    const float *pend;        // [2][40] rows left over by the previous call
    int n_pend;
    const void *rows;         // [n_rows][40] uint16 (scaled by 1/25.6 on load) or float32
    int n_rows;
    int rows_are_f32;
};

MWW_HD float nn_virtual_row(const NnInput &in, int vr, int f) {
    // vr < 0: ring history; then pending rows; then this call's rows
    if (vr < 0) return in.ring0[(2 + vr) * kNumChannels + f];
    if (vr < in.n_pend) return in.pend[vr * kNumChannels + f];
    const int r = vr - in.n_pend;
    if (in.rows_are_f32) return static_cast<const float *>(in.rows)[(long long)r * kNumChannels + f];
    return (float)static_cast<const uint16_t *>(in.rows)[(long long)r * kNumChannels + f] * kFeatureScale;   // inference.py:93-94
}

// ---- phase: load ring state into the shared activation buffers (once per call) ----
template <int L>
MWW_HD void nn_load_state_l(int tid, float *sm, const float *state) {
    constexpr NnLayerGeom g = kGeom[L];
    const float *src = state + kStateOff[L + 1];
    for (int e = tid; e < g.ring * g.cin; e += kNnThreads) {
        const int r = e / g.cin, c = e - r * g.cin;
        sm[g.off + c * g.ld + (g.hp - g.ring) + r] = src[e];
    }
}
MWW_HD void nn_load_state(int tid, float *sm, const float *state) {
    nn_load_state_l<0>(tid, sm, state); nn_load_state_l<1>(tid, sm, state); nn_load_state_l<2>(tid, sm, state);
    nn_load_state_l<3>(tid, sm, state); nn_load_state_l<4>(tid, sm, state);
}

// ---- phase: im2col the chunk's feature rows: featJ[j][f][t] = row(3*(t0+t) + j - 2) ----
MWW_HD void nn_load_features(int tid, float *sm, const NnInput &in, int step0, int n) {
    float *feat = sm + kXFloats + kDFloats;
    for (int e = tid; e < 5 * kNumChannels * kTT; e += kNnThreads) {
        const int t = e % kTT;
        const int f = (e / kTT) % kNumChannels;
        const int j = e / (kTT * kNumChannels);
        float v = 0.f;
        if (t < n) v = nn_virtual_row(in, 3 * (step0 + t) + j - 2, f);
        feat[(j * kNumChannels + f) * kTT + t] = v;
    }
}

// ---- phase: first conv (5x1, stride 3, 40 -> 32, no bias) + ReLU ----
MWW_HD void nn_first_conv(int tid, float *sm, const NnWeightsF32 &W) {
    const float *feat = sm + kXFloats + kDFloats;
    const int o = tid & 31, q = tid >> 5;          // 8 steps per thread
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k = 0; k < 5 * kNumChannels; ++k) {
        const float w = W.w0[k * 32 + o];
        const float *x = feat + k * kTT + 8 * q;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, x[i], acc[i]);
    }
    float *dst = sm + kGeom[0].off + o * kGeom[0].ld + kGeom[0].hp + 8 * q;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = acc[i] > 0.f ? acc[i] : 0.f;
}

// ---- phase: depthwise conv over (ring ++ chunk) for block L; output D[c][t] ----
template <int L>
MWW_HD void nn_depthwise(int tid, float *sm, const NnWeightsF32 &W) {
    constexpr NnLayerGeom g = kGeom[L];
    constexpr int TPC = kNnThreads / g.cin;        // threads per channel
    constexpr int TN = kTT / TPC;                  // steps per thread
    const int c = tid % g.cin, part = tid / g.cin;
    const int t0 = part * TN;
    float w[g.kmax];
#pragma unroll
    for (int j = 0; j < g.kmax; ++j) w[j] = W.dw_w[L][j * g.cin + c];
    const float bias = W.dw_b[L][c];
    float acc[TN];
#pragma unroll
    for (int i = 0; i < TN; ++i) acc[i] = 0.f;
    // output step t reads columns hp - (kmax-1) + t + j, j = 0..kmax-1 (oldest tap first, like the ring concat)
    const float *x = sm + g.off + c * g.ld + (g.hp - (g.kmax - 1)) + t0;
#pragma unroll
    for (int i = 0; i < TN + g.kmax - 1; ++i) {
        const float xv = x[i];
#pragma unroll
        for (int tt = 0; tt < TN; ++tt) {
            const int j = i - tt;
            if (j >= 0 && j < g.kmax) acc[tt] = fmaf(w[j], xv, acc[tt]);
        }
    }
    float *d = sm + kXFloats + c * kDLd + t0;
#pragma unroll
    for (int i = 0; i < TN; ++i) d[i] = acc[i] + bias;
}

// ---- phase: pointwise 1x1 (cin -> 64) + folded-BN bias + ReLU into the next ring buffer ----
template <int L>
MWW_HD void nn_pointwise(int tid, float *sm, const NnWeightsF32 &W) {
    constexpr int cin = kGeom[L].cin;
    constexpr NnLayerGeom gn = kGeom[L + 1];
    const int o = tid & 63, h = tid >> 6;          // 16 steps per thread
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const float *d = sm + kXFloats + 16 * h;
    for (int k = 0; k < cin; ++k) {
        const float w = W.pw_w[L][k * 64 + o];
        const float *x = d + k * kDLd;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(w, x[i], acc[i]);
    }
    const float bias = W.pw_b[L][o];
    float *dst = sm + gn.off + o * gn.ld + gn.hp + 16 * h;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float v = acc[i] + bias; dst[i] = v > 0.f ? v : 0.f; }
}

// ---- phase: head, part 1: per-channel 17-tap partial sums into D[c][t] ----
MWW_HD void nn_head_partial(int tid, float *sm, const NnWeightsF32 &W) {
    constexpr NnLayerGeom g = kGeom[4];
    const int c = tid & 63, h = tid >> 6;
    float w[17];
#pragma unroll
    for (int j = 0; j < 17; ++j) w[j] = W.head_w[j * 64 + c];
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const float *x = sm + g.off + c * g.ld + (g.hp - 16) + 16 * h;
#pragma unroll
    for (int i = 0; i < 16 + 16; ++i) {
        const float xv = x[i];
#pragma unroll
        for (int tt = 0; tt < 16; ++tt) {
            const int j = i - tt;
            if (j >= 0 && j < 17) acc[tt] = fmaf(w[j], xv, acc[tt]);
        }
    }
    float *d = sm + kXFloats + c * kDLd + 16 * h;
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = acc[i];
}

MWW_HD float nn_sigmoid(float x) {
#if defined(__CUDA_ARCH__)
    return 1.0f / (1.0f + expf(-x));
#else
    return 1.0f / (1.0f + expf(-x));
#endif
}

// ---- phase: head, part 2: reduce over channels, bias, sigmoid ----
MWW_HD void nn_head_finish(int tid, float *sm, const NnWeightsF32 &W, int n, float *probs_out /* step0-relative */, float *logits_out) {
    if (tid >= kTT || tid >= n) return;
    const float *d = sm + kXFloats + tid;
    float acc = 0.f;
    for (int c = 0; c < 64; ++c) acc += d[c * kDLd];
    const float logit = acc + W.head_b[0];
    if (logits_out) logits_out[tid] = logit;
    probs_out[tid] = nn_sigmoid(logit);
}

// ---- phase pair: slide the ring histories left by n steps (read -> barrier -> write) ----
constexpr int kShiftPerThread = 11;   // ceil(64*22 / 128)
template <int L>
MWW_HD void nn_shift_read_l(int tid, const float *sm, int n, float (&tmp)[kShiftPerThread]) {
    constexpr NnLayerGeom g = kGeom[L];
#pragma unroll
    for (int q = 0; q < kShiftPerThread; ++q) {
        const int e = tid + q * kNnThreads;
        if (e < g.ring * g.cin) {
            const int c = e / g.ring, r = e - c * g.ring;
            tmp[q] = sm[g.off + c * g.ld + (g.hp - g.ring) + n + r];
        }
    }
}
template <int L>
MWW_HD void nn_shift_write_l(int tid, float *sm, const float (&tmp)[kShiftPerThread]) {
    constexpr NnLayerGeom g = kGeom[L];
#pragma unroll
    for (int q = 0; q < kShiftPerThread; ++q) {
        const int e = tid + q * kNnThreads;
        if (e < g.ring * g.cin) {
            const int c = e / g.ring, r = e - c * g.ring;
            sm[g.off + c * g.ld + (g.hp - g.ring) + r] = tmp[q];
        }
    }
}
MWW_HD void nn_shift_read(int tid, const float *sm, int n, float (&tmp)[5][kShiftPerThread]) {
    nn_shift_read_l<0>(tid, sm, n, tmp[0]); nn_shift_read_l<1>(tid, sm, n, tmp[1]); nn_shift_read_l<2>(tid, sm, n, tmp[2]);
    nn_shift_read_l<3>(tid, sm, n, tmp[3]); nn_shift_read_l<4>(tid, sm, n, tmp[4]);
}
MWW_HD void nn_shift_write(int tid, float *sm, const float (&tmp)[5][kShiftPerThread]) {
    nn_shift_write_l<0>(tid, sm, tmp[0]); nn_shift_write_l<1>(tid, sm, tmp[1]); nn_shift_write_l<2>(tid, sm, tmp[2]);
    nn_shift_write_l<3>(tid, sm, tmp[3]); nn_shift_write_l<4>(tid, sm, tmp[4]);
}

// ---- phases: write the ring state, first-conv ring and pending rows back (once per call) ----
template <int L>
MWW_HD void nn_store_state_l(int tid, const float *sm, float *state) {
    constexpr NnLayerGeom g = kGeom[L];
    float *dst = state + kStateOff[L + 1];
    for (int e = tid; e < g.ring * g.cin; e += kNnThreads) {
        const int r = e / g.cin, c = e - r * g.cin;
        dst[e] = sm[g.off + c * g.ld + (g.hp - g.ring) + r];
    }
}
// The new first-conv ring (two rows preceding the next unconsumed row) and the new pending rows are
// gathered from the OLD ring/pending/rows first; a barrier separates this from nn_tail_write.
struct NnTail { float ring_new[2], pend_new[2]; };
MWW_HD void nn_tail_read(int tid, const NnInput &in, int n_steps, int n_virtual_rows, NnTail &t) {
    const int consumed = 3 * n_steps;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = tid + q * kNnThreads;
        t.ring_new[q] = 0.f; t.pend_new[q] = 0.f;
        if (e < 2 * kNumChannels) {
            const int r = e / kNumChannels, f = e - r * kNumChannels;
            t.ring_new[q] = nn_virtual_row(in, consumed - 2 + r, f);
            const int vr = consumed + r;
            t.pend_new[q] = vr < n_virtual_rows ? nn_virtual_row(in, vr, f) : 0.f;
        }
    }
}
MWW_HD void nn_tail_write(int tid, const float *sm, float *state, float *pend_out, const NnTail &t) {
    nn_store_state_l<0>(tid, sm, state); nn_store_state_l<1>(tid, sm, state); nn_store_state_l<2>(tid, sm, state);
    nn_store_state_l<3>(tid, sm, state); nn_store_state_l<4>(tid, sm, state);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = tid + q * kNnThreads;
        if (e < 2 * kNumChannels) { state[e] = t.ring_new[q]; pend_out[e] = t.pend_new[q]; }
    }
}

}  // namespace mww

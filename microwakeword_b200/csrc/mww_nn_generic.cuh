// mww_nn_generic.cuh -- streaming MixedNet for ANY geometry the reference's builder can emit with its default
// block structure (microwakeword/mixednet.py:278-386: strided first conv, N MixConv blocks = ring-buffered grouped
// depthwise conv + 1x1 pointwise with folded BatchNorm + ReLU, a (head_rows - 1)-row ring, Flatten, Dense(1), sigmoid),
// fp32 and TFLite-int8, with the ring semantics of microwakeword/layers/stream.py:581-595.
//
// The tensor-core kernels (mww_nn_dev.cuh, mww_nn_live.cuh and their int8 twins) are compiled for the okay_nabu geometry
// (notebooks/basic_training_notebook.ipynb:503-509) -- the configuration BASELINE.json measures.  Models trained with
// other --pointwise_filters / --mixconv_kernel_sizes / --first_conv_* / --stride flags take THIS path: the geometry is a
// run-time argument, one CTA owns one stream for the whole call and steps it exactly like the reference's interpreter
// (one model step = `stride` new feature rows), with every ring resident in shared memory as a circular buffer --
// loaded from HBM once per call and written back once, instead of the reference's whole-ring shift per step
// (stream.py:586-590).  CUDA cores only (the contractions of one stream-step are 1 x K x N: nothing for a tensor core),
// plain loads / stores, a barrier between phases: chosen so that tests/host_emul can execute the very same phase functions
// thread by thread on the CPU and compare them with the oracle for several architectures (tests/test_generic_arch.py).
//
// Not supported (mww_create reports MWW_EUNSUPPORTED): residual connections, repeat_in_block > 1, max-pool / spatial
// attention variants (mixednet.py flags that change the graph's topology), first_conv_kernel_size < stride.
#pragma once

#include "mww_nn_i8_dev.cuh"

namespace mww {

constexpr int kGenMaxBlocks = 8;
constexpr int kGenThreads = 128;
constexpr int kGenMaxChannels = 512;
constexpr int kGenMaxKernel = 64;

struct GenArch {
    int c0, k0, stride, n_blocks, head_rows, ring0;       // ring0 = k0 - stride rows kept by the first conv
    int cin[kGenMaxBlocks], cout[kGenMaxBlocks], kmax[kGenMaxBlocks];
    int c_last, max_c;
    // per-stream state (elements; float or int8), oldest row first, [row][channel]: first-conv ring, block rings, head ring
    int st_blk[kGenMaxBlocks], st_head, state_elems;       // the first-conv ring starts at 0
    int pend_cap;                                          // rows of the pending-row buffer: max(stride - 1, 1)
    // shared memory (4-byte elements): window [k0][40], a[max_c], d[max_c], circular rings [kmax][cin], head [head_rows][c_last]
    int sm_win, sm_a, sm_d, sm_blk[kGenMaxBlocks], sm_head, sm_elems;
    long long macs_per_step;
};

// arch tensor of the model container (microwakeword_b200/model_file.py): [c0, k0, stride, 40, n_blocks, head_rows] then
// per block [cout, n_groups, k_1..k_4].  Returns 0 or a negative reason code (-1 malformed, -2 unsupported size).
MWW_HD int gen_arch_from_tensor(const int32_t *a, int n_ints, GenArch *g) {
    if (n_ints < 6) return -1;
    GenArch A = GenArch();
    A.c0 = a[0]; A.k0 = a[1]; A.stride = a[2]; A.n_blocks = a[4]; A.head_rows = a[5];
    if (a[3] != kNumChannels || A.n_blocks < 1 || A.n_blocks > kGenMaxBlocks || n_ints < 6 + 6 * A.n_blocks) return -1;
    if (A.c0 < 1 || A.c0 > kGenMaxChannels || A.k0 < 1 || A.k0 > kGenMaxKernel || A.stride < 1 || A.stride > A.k0 ||
        A.head_rows < 1 || A.head_rows > kGenMaxKernel) return -2;
    A.ring0 = A.k0 - A.stride;
    A.pend_cap = A.stride > 1 ? A.stride - 1 : 1;
    int cin = A.c0, off = A.ring0 * kNumChannels, sm = 0;
    A.max_c = A.c0;
    A.sm_win = sm; sm += A.k0 * kNumChannels;
    long long macs = (long long)A.k0 * kNumChannels * A.c0;
    for (int i = 0; i < A.n_blocks; ++i) {
        const int32_t *e = a + 6 + 6 * i;
        const int cout = e[0], groups = e[1];
        if (cout < 1 || cout > kGenMaxChannels || groups < 1 || groups > 4) return -2;
        int kmax = 0;
        for (int j = 0; j < groups; ++j) { if (e[2 + j] < 1 || e[2 + j] > kGenMaxKernel) return -2; kmax = e[2 + j] > kmax ? e[2 + j] : kmax; }
        A.cin[i] = cin; A.cout[i] = cout; A.kmax[i] = kmax;
        A.st_blk[i] = off; off += (kmax - 1) * cin;
        // MACs as the reference executes them: each MixConv group only has its own taps (mixednet.py:132-136 split)
        int rest = cin;
        for (int j = 0; j < groups; ++j) { const int n = j == 0 ? cin / groups + (cin - cin / groups * groups) : cin / groups; macs += (long long)n * e[2 + j]; rest -= n; }
        (void)rest;
        macs += (long long)cin * cout;
        A.max_c = cout > A.max_c ? cout : A.max_c;
        cin = cout;
    }
    A.c_last = cin;
    A.st_head = off; off += (A.head_rows - 1) * cin;
    A.state_elems = off;
    macs += (long long)A.head_rows * cin;
    A.macs_per_step = macs;
    A.sm_a = sm; sm += A.max_c;
    A.sm_d = sm; sm += A.max_c;
    for (int i = 0; i < A.n_blocks; ++i) { A.sm_blk[i] = sm; sm += A.kmax[i] * A.cin[i]; }
    A.sm_head = sm; sm += A.head_rows * A.c_last;
    A.sm_elems = sm;
    *g = A;
    return 0;
}

struct GenWeightsF32 {
    const float *w0;                                   // [k0][40][c0]
    const float *dw_w[kGenMaxBlocks];                  // [kmax][cin], zero padded at the front for the shorter MixConv kernels
    const float *dw_b[kGenMaxBlocks];                  // [cin]
    const float *pw_w[kGenMaxBlocks];                  // [cin][cout]  BatchNorm folded
    const float *pw_b[kGenMaxBlocks];                  // [cout]
    const float *head_w;                               // [head_rows][c_last]
    const float *head_b;                               // [1]
};

struct GenWeightsI8 {
    const int8_t *w0; const int32_t *b0, *m0, *s0;
    const int8_t *dw_w[kGenMaxBlocks]; const int32_t *dw_b[kGenMaxBlocks], *dw_m[kGenMaxBlocks], *dw_s[kGenMaxBlocks];
    const int8_t *pw_w[kGenMaxBlocks]; const int32_t *pw_b[kGenMaxBlocks], *pw_m[kGenMaxBlocks], *pw_s[kGenMaxBlocks];
    const int8_t *head_w; const int8_t *lut;
    int32_t head_bias, head_mult, head_shift;
    int32_t zp[4 + 2 * kGenMaxBlocks];   // [in, first-conv out, (depthwise out, pointwise out) per block, logit, prob]
    float in_scale;
};

// this call's input of one stream: the virtual row sequence  V = first-conv ring ++ pending rows ++ rows
template <typename T>
struct GenInput {
    const T *state;            // this stream's state (the first-conv ring is its first ring0 rows)
    const T *pend;             // [pend_cap][40]
    int n_pend;
    const void *rows;          // [n_rows][40] uint16 / float32 (/ int8 for a quantised model)
    int n_rows;
    int row_type;              // 0 uint16, 1 float32, 2 int8
};

// ------------------------------------------------------------------------------------------------------------------
// fp32

MWW_HD float gen_virtual_f32(const GenArch &A, const GenInput<float> &in, int v, int f) {
    if (v < A.ring0) return in.state[v * kNumChannels + f];
    v -= A.ring0;
    if (v < in.n_pend) return in.pend[v * kNumChannels + f];
    const long long e = (long long)(v - in.n_pend) * kNumChannels + f;
    if (in.row_type == 1) return static_cast<const float *>(in.rows)[e];
    return (float)static_cast<const uint16_t *>(in.rows)[e] * kFeatureScale;                 // inference.py:93-94
}

// circular ring of block b / the head: slot 0 is the next write position after loading, slots 1.. hold oldest..newest
template <typename T, typename S>
MWW_HD void gen_load_ring(int tid, S *buf, const T *src, int rows_kept, int c) {
    for (int e = tid; e < c; e += kGenThreads) buf[e] = S(0);
    for (int e = tid; e < rows_kept * c; e += kGenThreads) buf[c + e] = (S)src[e];
}
// after the call: `pos` is the next write slot = the oldest row; the kept rows are the kmax - 1 newest, oldest first
template <typename T, typename S>
MWW_HD void gen_store_ring(int tid, const S *buf, T *dst, int slots, int c, int pos) {
    for (int e = tid; e < (slots - 1) * c; e += kGenThreads) {
        const int r = e / c, ch = e - r * c;
        int slot = pos + 1 + r;
        slot = slot >= slots ? slot - slots : slot;
        dst[e] = (T)buf[slot * c + ch];
    }
}

MWW_HD void gen_f32_load_state(int tid, float *sm, const GenArch &A, const float *state) {
    for (int b = 0; b < A.n_blocks; ++b) gen_load_ring(tid, sm + A.sm_blk[b], state + A.st_blk[b], A.kmax[b] - 1, A.cin[b]);
    gen_load_ring(tid, sm + A.sm_head, state + A.st_head, A.head_rows - 1, A.c_last);
}

// phase: the first conv's window of step t = V[stride t .. stride t + k0)
MWW_HD void gen_f32_window(int tid, float *sm, const GenArch &A, const GenInput<float> &in, int t) {
    float *win = sm + A.sm_win;
    for (int e = tid; e < A.k0 * kNumChannels; e += kGenThreads) {
        const int k = e / kNumChannels, f = e - k * kNumChannels;
        win[e] = gen_virtual_f32(A, in, A.stride * t + k, f);
    }
}
// phase: first conv + ReLU (mixednet.py:317-331) -> a[c0]
MWW_HD void gen_f32_first_conv(int tid, float *sm, const GenArch &A, const GenWeightsF32 &W) {
    const float *win = sm + A.sm_win;
    float *a = sm + A.sm_a;
    const int K = A.k0 * kNumChannels;
    for (int o = tid; o < A.c0; o += kGenThreads) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(win[k], W.w0[k * A.c0 + o], acc);
        a[o] = acc > 0.f ? acc : 0.f;
    }
}
// phase: the new row enters block b's ring (stream.py:584-590) and the grouped depthwise conv runs over the last kmax rows
MWW_HD void gen_f32_depthwise(int tid, float *sm, const GenArch &A, const GenWeightsF32 &W, int b, int pos) {
    const int cin = A.cin[b], kmax = A.kmax[b];
    float *buf = sm + A.sm_blk[b];
    const float *a = sm + A.sm_a;
    float *d = sm + A.sm_d;
    for (int c = tid; c < cin; c += kGenThreads) {
        buf[pos * cin + c] = a[c];
        float acc = 0.f;
        int slot = pos + 1;
        for (int j = 0; j < kmax; ++j) {
            slot = slot >= kmax ? slot - kmax : slot;
            acc = fmaf(buf[slot * cin + c], W.dw_w[b][j * cin + c], acc);
            ++slot;
        }
        d[c] = acc + W.dw_b[b][c];
    }
}
// phase: 1x1 projection (BatchNorm folded) + ReLU -> a[cout]
MWW_HD void gen_f32_pointwise(int tid, float *sm, const GenArch &A, const GenWeightsF32 &W, int b) {
    const int cin = A.cin[b], cout = A.cout[b];
    const float *d = sm + A.sm_d;
    float *a = sm + A.sm_a;
    for (int o = tid; o < cout; o += kGenThreads) {
        float acc = 0.f;
        for (int c = 0; c < cin; ++c) acc = fmaf(d[c], W.pw_w[b][c * cout + o], acc);
        acc += W.pw_b[b][o];
        a[o] = acc > 0.f ? acc : 0.f;
    }
}
// phase: head, part 1: the new row enters the head ring; per-channel partial dot products over head_rows rows -> d[c]
MWW_HD void gen_f32_head_partial(int tid, float *sm, const GenArch &A, const GenWeightsF32 &W, int pos) {
    const int cl = A.c_last, hr = A.head_rows;
    float *buf = sm + A.sm_head;
    const float *a = sm + A.sm_a;
    float *d = sm + A.sm_d;
    for (int c = tid; c < cl; c += kGenThreads) {
        buf[pos * cl + c] = a[c];
        float acc = 0.f;
        int slot = pos + 1;
        for (int j = 0; j < hr; ++j) {
            slot = slot >= hr ? slot - hr : slot;
            acc = fmaf(buf[slot * cl + c], W.head_w[j * cl + c], acc);
            ++slot;
        }
        d[c] = acc;
    }
}
// phase: head, part 2 (thread 0): reduce over channels, bias, sigmoid (mixednet.py:383-384)
MWW_HD void gen_f32_head_finish(int tid, const float *sm, const GenArch &A, const GenWeightsF32 &W, float *prob_out) {
    if (tid != 0) return;
    const float *d = sm + A.sm_d;
    float acc = 0.f;
    for (int c = 0; c < A.c_last; ++c) acc += d[c];
    *prob_out = nn_sigmoid(acc + W.head_b[0]);
}
// tail, part 1: the rows the NEXT call starts from -- new first-conv ring = V[sT .. sT + ring0), new pending rows after it --
// are gathered into the (now idle) window buffer while the old ring / pending rows are still intact
MWW_HD void gen_f32_tail_gather(int tid, float *sm, const GenArch &A, const GenInput<float> &in, int n_steps) {
    float *win = sm + A.sm_win;
    const int n_virtual = A.ring0 + in.n_pend + in.n_rows, first = A.stride * n_steps;
    const int n_keep = n_virtual - first;                                   // ring0 + new pending rows  (< k0)
    for (int e = tid; e < n_keep * kNumChannels; e += kGenThreads) {
        const int k = e / kNumChannels, f = e - k * kNumChannels;
        win[e] = gen_virtual_f32(A, in, first + k, f);
    }
}
// tail, part 2 (after a barrier): write everything back
MWW_HD void gen_f32_tail_store(int tid, const float *sm, const GenArch &A, float *state, float *pend, int n_keep, const int *pos /* [n_blocks + 1] */) {
    const float *win = sm + A.sm_win;
    for (int e = tid; e < A.ring0 * kNumChannels; e += kGenThreads) state[e] = win[e];
    for (int e = tid; e < (n_keep - A.ring0) * kNumChannels; e += kGenThreads) pend[e] = win[A.ring0 * kNumChannels + e];
    for (int b = 0; b < A.n_blocks; ++b) gen_store_ring(tid, sm + A.sm_blk[b], state + A.st_blk[b], A.kmax[b], A.cin[b], pos[b]);
    gen_store_ring(tid, sm + A.sm_head, state + A.st_head, A.head_rows, A.c_last, pos[A.n_blocks]);
}

// ------------------------------------------------------------------------------------------------------------------
// int8 (TFLite reference integer kernels; same arithmetic helpers as mww_nn_i8_dev.cuh).
// Activations live in shared memory as their raw quantised value in an int32.

MWW_HD int32_t gen_virtual_i8(const GenArch &A, const GenWeightsI8 &W, const GenInput<int8_t> &in, int v, int f) {
    if (v < A.ring0) return in.state[v * kNumChannels + f];
    v -= A.ring0;
    if (v < in.n_pend) return in.pend[v * kNumChannels + f];
    const long long e = (long long)(v - in.n_pend) * kNumChannels + f;
    if (in.row_type == 2) return static_cast<const int8_t *>(in.rows)[e];
    const float x = in.row_type == 1 ? static_cast<const float *>(in.rows)[e]
                                     : (float)static_cast<const uint16_t *>(in.rows)[e] * kFeatureScale;
    return nnq_quantize(x, W.in_scale, W.zp[0]);                                              // inference.py:127-147
}
MWW_HD void gen_i8_load_state(int tid, int32_t *sm, const GenArch &A, const int8_t *state) {
    for (int b = 0; b < A.n_blocks; ++b) gen_load_ring(tid, sm + A.sm_blk[b], state + A.st_blk[b], A.kmax[b] - 1, A.cin[b]);
    gen_load_ring(tid, sm + A.sm_head, state + A.st_head, A.head_rows - 1, A.c_last);
}
MWW_HD void gen_i8_window(int tid, int32_t *sm, const GenArch &A, const GenWeightsI8 &W, const GenInput<int8_t> &in, int t) {
    int32_t *win = sm + A.sm_win;
    for (int e = tid; e < A.k0 * kNumChannels; e += kGenThreads) {
        const int k = e / kNumChannels, f = e - k * kNumChannels;
        win[e] = gen_virtual_i8(A, W, in, A.stride * t + k, f);
    }
}
MWW_HD void gen_i8_first_conv(int tid, int32_t *sm, const GenArch &A, const GenWeightsI8 &W) {
    const int32_t *win = sm + A.sm_win;
    int32_t *a = sm + A.sm_a;
    const int K = A.k0 * kNumChannels;
    const int32_t zp_in = W.zp[0], zp_out = W.zp[1];
    for (int o = tid; o < A.c0; o += kGenThreads) {
        int32_t acc = 0;
        for (int k = 0; k < K; ++k) acc += (win[k] - zp_in) * (int32_t)W.w0[k * A.c0 + o];
        acc += W.b0[o];
        a[o] = requant_rel(acc, W.m0[o], W.s0[o], zp_out, true) + zp_out;
    }
}
MWW_HD void gen_i8_depthwise(int tid, int32_t *sm, const GenArch &A, const GenWeightsI8 &W, int b, int pos) {
    const int cin = A.cin[b], kmax = A.kmax[b];
    int32_t *buf = sm + A.sm_blk[b];
    const int32_t *a = sm + A.sm_a;
    int32_t *d = sm + A.sm_d;
    const int32_t zp_in = W.zp[1 + 2 * b], zp_d = W.zp[2 + 2 * b];
    for (int c = tid; c < cin; c += kGenThreads) {
        buf[pos * cin + c] = a[c];
        int32_t acc = 0;
        int slot = pos + 1;
        for (int j = 0; j < kmax; ++j) {
            slot = slot >= kmax ? slot - kmax : slot;
            acc += (buf[slot * cin + c] - zp_in) * (int32_t)W.dw_w[b][j * cin + c];
            ++slot;
        }
        acc += W.dw_b[b][c];
        d[c] = requant_rel(acc, W.dw_m[b][c], W.dw_s[b][c], zp_d, false) + zp_d;
    }
}
MWW_HD void gen_i8_pointwise(int tid, int32_t *sm, const GenArch &A, const GenWeightsI8 &W, int b) {
    const int cin = A.cin[b], cout = A.cout[b];
    const int32_t *d = sm + A.sm_d;
    int32_t *a = sm + A.sm_a;
    const int32_t zp_d = W.zp[2 + 2 * b], zp_p = W.zp[3 + 2 * b];
    for (int o = tid; o < cout; o += kGenThreads) {
        int32_t acc = 0;
        for (int c = 0; c < cin; ++c) acc += (d[c] - zp_d) * (int32_t)W.pw_w[b][c * cout + o];
        acc += W.pw_b[b][o];
        a[o] = requant_rel(acc, W.pw_m[b][o], W.pw_s[b][o], zp_p, true) + zp_p;
    }
}
MWW_HD void gen_i8_head_partial(int tid, int32_t *sm, const GenArch &A, const GenWeightsI8 &W, int pos) {
    const int cl = A.c_last, hr = A.head_rows;
    int32_t *buf = sm + A.sm_head;
    const int32_t *a = sm + A.sm_a;
    int32_t *d = sm + A.sm_d;
    const int32_t zp_in = W.zp[1 + 2 * A.n_blocks];
    for (int c = tid; c < cl; c += kGenThreads) {
        buf[pos * cl + c] = a[c];
        int32_t acc = 0;
        int slot = pos + 1;
        for (int j = 0; j < hr; ++j) {
            slot = slot >= hr ? slot - hr : slot;
            acc += (buf[slot * cl + c] - zp_in) * (int32_t)W.head_w[j * cl + c];
            ++slot;
        }
        d[c] = acc;
    }
}
// FULLY_CONNECTED requant -> LOGISTIC LUT -> QUANTIZE to uint8 -> Model.dequantize_output_data (/255, inference.py:162-170)
MWW_HD void gen_i8_head_finish(int tid, const int32_t *sm, const GenArch &A, const GenWeightsI8 &W, float *prob_out) {
    if (tid != 0) return;
    const int32_t *d = sm + A.sm_d;
    int32_t acc = 0;
    for (int c = 0; c < A.c_last; ++c) acc += d[c];
    acc += W.head_bias;
    const int32_t zp_fc = W.zp[2 + 2 * A.n_blocks];
    const int32_t logit = requant_rel(acc, W.head_mult, W.head_shift, zp_fc, false) + zp_fc;
    const int out_u8 = (int)W.lut[(uint8_t)(int8_t)logit] + 128;
    *prob_out = (1.0f / 255.0f) * (float)out_u8;
}
MWW_HD void gen_i8_tail_gather(int tid, int32_t *sm, const GenArch &A, const GenWeightsI8 &W, const GenInput<int8_t> &in, int n_steps) {
    int32_t *win = sm + A.sm_win;
    const int n_virtual = A.ring0 + in.n_pend + in.n_rows, first = A.stride * n_steps;
    const int n_keep = n_virtual - first;
    for (int e = tid; e < n_keep * kNumChannels; e += kGenThreads) {
        const int k = e / kNumChannels, f = e - k * kNumChannels;
        win[e] = gen_virtual_i8(A, W, in, first + k, f);
    }
}
MWW_HD void gen_i8_tail_store(int tid, const int32_t *sm, const GenArch &A, int8_t *state, int8_t *pend, int n_keep, const int *pos) {
    const int32_t *win = sm + A.sm_win;
    for (int e = tid; e < A.ring0 * kNumChannels; e += kGenThreads) state[e] = (int8_t)win[e];
    for (int e = tid; e < (n_keep - A.ring0) * kNumChannels; e += kGenThreads) pend[e] = (int8_t)win[A.ring0 * kNumChannels + e];
    for (int b = 0; b < A.n_blocks; ++b) gen_store_ring(tid, sm + A.sm_blk[b], state + A.st_blk[b], A.kmax[b], A.cin[b], pos[b]);
    gen_store_ring(tid, sm + A.sm_head, state + A.st_head, A.head_rows, A.c_last, pos[A.n_blocks]);
}
// reset value of one element of the state / pending buffers: the zero point of the tensor the ring buffers
// (quantised state variables hold real 0, utils.py:333)
MWW_HD int8_t gen_i8_reset_value(const GenArch &A, const GenWeightsI8 &W, int e /* index into the state */) {
    if (e < A.ring0 * kNumChannels) return (int8_t)W.zp[0];
    for (int b = A.n_blocks - 1; b >= 0; --b)
        if (e >= A.st_blk[b] && e < A.st_blk[b] + (A.kmax[b] - 1) * A.cin[b]) return (int8_t)W.zp[1 + 2 * b];
    return (int8_t)W.zp[1 + 2 * A.n_blocks];
}

}  // namespace mww

// mww_nn_i8_live.cuh -- int8 MixedNet, LIVE-STEP formulation: one model step (three new 10 ms feature rows) for many
// streams per launch, bit-exact TFLite int8 semantics (same arithmetic as mww_nn_i8_dev.cuh).
//
// This is the shape the reference's deployed models actually run in: an int8 streaming .tflite stepped once per
// 30 ms (microwakeword/inference.py:109-123).  The ring state is 4 176 BYTES per stream here, so a step moves
// ~4.7 KB per stream through HBM (SURVEY.md 8d "live step mode, int8 state") against 18.4 KB for the fp32 model.
//
// Structure mirrors mww_nn_live.cuh: a CTA owns 32 streams per group (the M dimension of the IMMA contractions:
// first conv K = 200 -> 224, 1x1 projections K = 32 / 64), all MMA weights stay resident in shared memory, rings are
// kept ROTATED between live calls (LiveHeads) so only the new row of each ring is written, and every activation
// lives in shared memory as raw int8 bytes [stream][channel] so the depthwise output is directly the next IMMA's A
// operand.  A thread of the depthwise / head stages owns 4 consecutive channels (one 32-bit word per ring row).
#pragma once

#include "mww_nn_i8_dev.cuh"
#include "mww_nn_live.cuh"

namespace mww {

constexpr int kLiveQThreads = 256;
// shared memory (bytes)
constexpr int kLqOffW0 = 0;                                   // [32 n][kW0Pitch]
constexpr int kLqOffPw = kLqOffW0 + 32 * kW0Pitch;            // 4 x [64 n][kPwPitch]
constexpr int kLqOffA = kLqOffPw + 4 * 64 * kPwPitch;         // first-conv window [32 streams][kW0Pitch]; later the head's partial sums
constexpr int kLqOffH = kLqOffA + kLiveStreams * kW0Pitch;    // newest activation row [32 streams][kPwPitch]
constexpr int kLqOffD = kLqOffH + kLiveStreams * kPwPitch;    // depthwise output      [32 streams][kPwPitch]
constexpr int kLiveQSmemBytes = kLqOffD + kLiveStreams * kPwPitch;       // 40.9 KB -> 4 CTAs / SM at <= 64 registers

struct LiveInputI8 {
    const int8_t *state;       // [S][4176]
    const int8_t *pend;        // [S][2][40]
    int n_pend;
    const void *rows;          // [S][3][40] uint16 / float32 / int8
    long long rows_stream_stride_bytes;
    int row_type;              // 0 uint16, 1 float32, 2 int8
};

struct alignas(16) LiveF4 { float v[4]; };
MWW_HD int32_t livq_sext(uint32_t w, int q) { return (int32_t)(int8_t)(w >> (8 * q)); }
MWW_HD uint32_t livq_pack(int32_t a, int32_t b, int32_t c, int32_t d) {
    return (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)(d & 0xFF) << 24);
}
MWW_HD uint32_t livq_ld32(const int8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }

// four consecutive input features (feature-row word e4 of this call's 30) as raw int8, quantised like
// Model.quantize_input_data (inference.py:127-147) when the rows are uint16 / float32
MWW_HD uint32_t livq_row_word(const LiveInputI8 &in, const NnWeightsI8 &W, size_t su, unsigned e4) {
    const char *base = static_cast<const char *>(in.rows) + su * (size_t)in.rows_stream_stride_bytes;
    if (in.row_type == 2) return reinterpret_cast<const uint32_t *>(base)[e4];
    float x[4];
    if (in.row_type == 1) {
        const LiveF4 v = *reinterpret_cast<const LiveF4 *>(base + 16 * (size_t)e4);
        x[0] = v.v[0]; x[1] = v.v[1]; x[2] = v.v[2]; x[3] = v.v[3];
    } else {
        struct alignas(8) U16x4 { uint16_t v[4]; };
        const U16x4 v = *reinterpret_cast<const U16x4 *>(base + 8 * (size_t)e4);
        if (W.qlut) return livq_pack(W.qlut[v.v[0]], W.qlut[v.v[1]], W.qlut[v.v[2]], W.qlut[v.v[3]]);
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = (float)v.v[q] * kFeatureScale;
    }
    return livq_pack(nnq_quantize(x[0], W.in_scale, W.zp[0]), nnq_quantize(x[1], W.in_scale, W.zp[0]),
                     nnq_quantize(x[2], W.in_scale, W.zp[0]), nnq_quantize(x[3], W.in_scale, W.zp[0]));
}

// ---- once per CTA: all MMA weights (28 KB) ----
MWW_HD void livq_load_weights(int tid, uint8_t *smb, const NnWeightsI8 &W) {
    struct alignas(16) Vec16 { uint32_t v[4]; };
    for (int e = tid; e < 32 * kW0Pitch / 16; e += kLiveQThreads)
        reinterpret_cast<Vec16 *>(smb + kLqOffW0)[e] = reinterpret_cast<const Vec16 *>(W.w0t)[e];
    for (int L = 0; L < 4; ++L)
        for (int e = tid; e < 64 * kPwPitch / 16; e += kLiveQThreads)
            reinterpret_cast<Vec16 *>(smb + kLqOffPw + L * 64 * kPwPitch)[e] = reinterpret_cast<const Vec16 *>(W.pwt[L])[e];
}

// ---- phase: first-conv window.  Per stream it is the concatenation (in 32-bit words of 4 int8)
//     window[0:50] = state[0:20] (2-row first-conv ring) ++ pend[0:10 p] ++ rows[0:30 - 10 p],      p = n_pend
// new ring = window[30:50], new pend = rows[30 - 10 p : 30] (unused pending slots hold the input zero point).
// A warp owns 4 streams; lane l builds words l and 32 + l.  Loads are unconditional on clamped addresses.
constexpr int kLqKeep = 3;
MWW_HD void livq_build_a(int tid, uint8_t *smb, const LiveInputI8 &in, const NnWeightsI8 &W, long long s0, int n_valid,
                         uint32_t (&keep)[4][kLqKeep]) {
    const int warp = tid >> 5;
    const unsigned lane = (unsigned)(tid & 31), np10 = (unsigned)in.n_pend * 10u;
    const uint32_t zpw = livq_pack(W.zp[0], W.zp[0], W.zp[0], W.zp[0]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sl = warp * 4 + q;
        const bool ok = sl < n_valid;
        const size_t su = (size_t)(s0 + (ok ? sl : 0));
        const int8_t *st = in.state + su * (size_t)kStateFloats;
        const int8_t *pd = in.pend + su * (size_t)(2 * kNumChannels);
        uint32_t *a = reinterpret_cast<uint32_t *>(smb + kLqOffA + sl * kW0Pitch);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const unsigned wd = lane + 32u * (unsigned)i, wc = wd < 49u ? wd : 49u;
            const bool in_ring = wc < 20u, in_pend = !in_ring && wc < 20u + np10, is_row = !in_ring && !in_pend;
            const uint32_t sv = livq_ld32(in_ring ? st + 4u * wc : (in_pend ? pd + 4u * (wc - 20u) : st));
            const uint32_t rv = livq_row_word(in, W, su, is_row ? wc - 20u - np10 : 0u);
            uint32_t v = is_row ? rv : sv;
            v = (ok && wd < 50u) ? v : 0u;
            keep[q][i] = v;
            if (wd < (unsigned)(kW0Pitch / 4)) a[wd] = v;                       // words 50..59 (k = 200..239) are zero padding
        }
        const bool live = ok && lane < np10;
        const uint32_t pv = livq_row_word(in, W, su, live ? 30u - np10 + lane : 0u);
        keep[q][2] = live ? pv : zpw;
    }
}
MWW_HD void livq_write_tail(int tid, int8_t *state, int8_t *pend, long long s0, int n_valid, const uint32_t (&keep)[4][kLqKeep]) {
    const int warp = tid >> 5;
    const unsigned lane = (unsigned)(tid & 31);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int sl = warp * 4 + q;
        if (sl >= n_valid) continue;
        const size_t su = (size_t)(s0 + sl);
        uint32_t *st = reinterpret_cast<uint32_t *>(state + su * (size_t)kStateFloats);
        if (lane >= 30u) st[lane - 30u] = keep[q][0];                            // window words 30, 31
        if (lane < 18u) st[lane + 2u] = keep[q][1];                              // window words 32 .. 49
        if (lane < 20u) reinterpret_cast<uint32_t *>(pend + su * (size_t)(2 * kNumChannels))[lane] = keep[q][2];
    }
}

// ---- epilogues: requantise an IMMA tile into H8[stream][channel] (raw int8, ReLU clamps at the zero point) ----
MWW_HD void livq_fc_store_tile(uint8_t *smb, const NnWeightsI8 &W, int r0, int n0, int lane, const int32_t (&c)[4]) {
    const int g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = r0 + g + ((i & 2) ? 8 : 0), o = n0 + 2 * tig + (i & 1);
        smb[kLqOffH + row * kPwPitch + o] = (uint8_t)(int8_t)(requant_rel(c[i] + W.b0f[o], W.m0[o], W.s0[o], W.zp[1], true) + W.zp[1]);
    }
}
template <int L>
MWW_HD void livq_pw_store_tile(uint8_t *smb, const NnWeightsI8 &W, int r0, int n0, int lane, const int32_t (&c)[4]) {
    const int g = lane >> 2, tig = lane & 3;
    const int32_t zp = W.zp[3 + 2 * L];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = r0 + g + ((i & 2) ? 8 : 0), o = n0 + 2 * tig + (i & 1);
        smb[kLqOffH + row * kPwPitch + o] = (uint8_t)(int8_t)(requant_rel(c[i] + W.pw_bf[L][o], W.pw_m[L][o], W.pw_s[L][o], zp, true) + zp);
    }
}

// ---- depthwise of block L: ring rows stream through registers as 32-bit words of 4 channels ----
// thread -> 4-channel group c4 = tid % (cin / 4), stream subgroup = tid / (cin / 4); NS = 1 (cin 32) or 2 (cin 64)
// streams per thread.  Taps are looked up rotated (physical row p holds logical row (p - head) mod R).
template <int R, int CIN>
MWW_HD void livq_ring_accumulate(int tid, const int8_t *taps, int32_t zp_in, const int8_t *state_ring, long long s0, int n_valid, int head,
                                 const uint8_t *h8, int32_t (&acc)[kLiveStreams * CIN / 4 / kLiveQThreads][4],
                                 uint32_t (&hw)[kLiveStreams * CIN / 4 / kLiveQThreads]) {
    constexpr int G = CIN / 4, NS = kLiveStreams * G / kLiveQThreads;
    const int c4 = tid % G, sub = tid / G;
    uint32_t x[NS][R];
#pragma unroll
    for (int u = 0; u < NS; ++u) {
        const int sl = sub * NS + u;
        const int8_t *ring = state_ring + (size_t)(s0 + (sl < n_valid ? sl : 0)) * kStateFloats + 4 * c4;
#pragma unroll
        for (int r = 0; r < R; ++r) x[u][r] = livq_ld32(ring + r * CIN);
        hw[u] = *reinterpret_cast<const uint32_t *>(h8 + sl * kPwPitch + 4 * c4);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] = 0;
    }
#pragma unroll
    for (int p = 0; p < R; ++p) {
        const int j = p - head < 0 ? p - head + R : p - head;
        const uint32_t w = livq_ld32(taps + j * CIN + 4 * c4);
#pragma unroll
        for (int u = 0; u < NS; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[u][q] += (livq_sext(x[u][p], q) - zp_in) * livq_sext(w, q);
    }
    const uint32_t wn = livq_ld32(taps + R * CIN + 4 * c4);                     // tap of the newest row
#pragma unroll
    for (int u = 0; u < NS; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[u][q] += (livq_sext(hw[u], q) - zp_in) * livq_sext(wn, q);
}

template <int L>
MWW_HD void livq_depthwise(int tid, uint8_t *smb, const NnWeightsI8 &W, int8_t *state, long long s0, int n_valid, int head) {
    constexpr NnLayerGeom g = kGeom[L];
    constexpr int G = g.cin / 4, NS = kLiveStreams * G / kLiveQThreads;
    constexpr int ring_off = kStateOff[L + 1];
    const int c4 = tid % G, sub = tid / G;
    int32_t acc[NS][4];
    uint32_t hw[NS];
    livq_ring_accumulate<g.ring, g.cin>(tid, W.dw_w[L], nnq_ring_zp(W, L), state + ring_off, s0, n_valid, head, smb + kLqOffH, acc, hw);
    const int32_t zp_out = W.zp[2 + 2 * L];
#pragma unroll
    for (int u = 0; u < NS; ++u) {
        const int sl = sub * NS + u;
        int32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 4 * c4 + q;
            o[q] = requant_rel(acc[u][q] + W.dw_b[L][c], W.dw_m[L][c], W.dw_s[L][c], zp_out, false) + zp_out;
        }
        *reinterpret_cast<uint32_t *>(smb + kLqOffD + sl * kPwPitch + 4 * c4) = livq_pack(o[0], o[1], o[2], o[3]);
        if (sl < n_valid)
            *reinterpret_cast<uint32_t *>(state + (size_t)(s0 + sl) * kStateFloats + ring_off + head * g.cin + 4 * c4) = hw[u];   // oldest row <- newest
    }
}

// ---- head: 17-tap dot per (stream, 4 channels) -> int32 partial sums in the (now free) window buffer ----
MWW_HD void livq_head_partial(int tid, uint8_t *smb, const NnWeightsI8 &W, int8_t *state, long long s0, int n_valid, int head) {
    constexpr int ring_off = kStateOff[5];
    const int c4 = tid % 16, sub = tid / 16;
    int32_t acc[2][4];
    uint32_t hw[2];
    livq_ring_accumulate<16, 64>(tid, W.head_w, W.zp[9], state + ring_off, s0, n_valid, head, smb + kLqOffH, acc, hw);
    int32_t *scratch = reinterpret_cast<int32_t *>(smb + kLqOffA);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int sl = sub * 2 + u;
        scratch[sl * 16 + c4] = acc[u][0] + acc[u][1] + acc[u][2] + acc[u][3];
        if (sl < n_valid) *reinterpret_cast<uint32_t *>(state + (size_t)(s0 + sl) * kStateFloats + ring_off + head * 64 + 4 * c4) = hw[u];
    }
}
// FULLY_CONNECTED requant -> LOGISTIC LUT -> QUANTIZE to uint8 -> Model.dequantize_output_data (/255)
MWW_HD void livq_head_finish(int tid, const uint8_t *smb, const NnWeightsI8 &W, long long s0, int n_valid, float *probs, long long probs_stride) {
    if (tid >= kLiveStreams || tid >= n_valid) return;
    const int32_t *scratch = reinterpret_cast<const int32_t *>(smb + kLqOffA) + tid * 16;
    int32_t acc = 0;
    for (int i = 0; i < 16; ++i) acc += scratch[i];
    const int32_t logit = requant_rel(acc + W.head_bias, W.head_mult, W.head_shift, W.zp[10], false) + W.zp[10];
    const int32_t out_u8 = (int32_t)W.lut[(uint8_t)(int8_t)logit] + 128;
    probs[(s0 + tid) * probs_stride] = (1.0f / 255.0f) * (float)out_u8;          // inference.py:162-170
}

// ---- back to the canonical layout (int8 rings; same column walk as live_canonicalise_column) ----
MWW_HD void livq_canonicalise_column(int8_t *state, long long s, int col, const LiveHeads &heads) {
    int i = 0, c = col;
    while (c >= live_ring_cols(i)) { c -= live_ring_cols(i); ++i; }
    const int R = live_ring_rows(i), C = live_ring_cols(i), h = heads.h[i];
    if (h == 0) return;
    int off = 2 * kNumChannels;
    for (int k = 0; k < i; ++k) off += live_ring_rows(k) * live_ring_cols(k);
    int8_t *ring = state + (size_t)s * kStateFloats + off + c;
    int8_t tmp[22];
    for (int j = 0; j < R; ++j) { const int p = h + j >= R ? h + j - R : h + j; tmp[j] = ring[p * C]; }
    for (int j = 0; j < R; ++j) ring[j * C] = tmp[j];
}

#if defined(__CUDACC__)
// first conv: 8 warps = 2 stream tiles x 4 channel tiles, 7 k-steps of 32
MWW_D void livq_first_conv_mma(int tid, uint8_t *smb, const NnWeightsI8 &W) {
    const int warp = tid >> 5, lane = tid & 31;
    const int r0 = 16 * (warp >> 2), n0 = 8 * (warp & 3);
    const int8_t *a8 = reinterpret_cast<const int8_t *>(smb + kLqOffA), *w0 = reinterpret_cast<const int8_t *>(smb + kLqOffW0);
    int32_t c[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
        FragA8 a;
        FragB8 b;
        load_frag_a8(a8, kW0Pitch, 32 * ks, r0, lane, a);
        load_frag_b8(w0, kW0Pitch, 32 * ks, n0, lane, b);
        mma_s8(c, a, b);
    }
    livq_fc_store_tile(smb, W, r0, n0, lane, c);
}
// 1x1 of block L: 8 warps = 2 stream tiles x 4 pairs of channel tiles
template <int L>
MWW_D void livq_pointwise_mma(int tid, uint8_t *smb, const NnWeightsI8 &W) {
    constexpr int cin = kGeom[L].cin;
    const int warp = tid >> 5, lane = tid & 31;
    const int r0 = 16 * (warp >> 2), n0 = 16 * (warp & 3);
    const int8_t *d8 = reinterpret_cast<const int8_t *>(smb + kLqOffD), *wt = reinterpret_cast<const int8_t *>(smb + kLqOffPw + L * 64 * kPwPitch);
    int32_t c[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
    for (int ks = 0; ks < cin / 32; ++ks) {
        FragA8 a;
        FragB8 b0, b1;
        load_frag_a8(d8, kPwPitch, 32 * ks, r0, lane, a);
        load_frag_b8(wt, kPwPitch, 32 * ks, n0, lane, b0);
        load_frag_b8(wt, kPwPitch, 32 * ks, n0 + 8, lane, b1);
        mma_s8(c[0], a, b0);
        mma_s8(c[1], a, b1);
    }
    livq_pw_store_tile<L>(smb, W, r0, n0, lane, c[0]);
    livq_pw_store_tile<L>(smb, W, r0, n0 + 8, lane, c[1]);
}
#endif

}  // namespace mww

// tests/host_emul/emul_nn.cc -- TEST INFRASTRUCTURE: CPU execution of the product's fp32 MixedNet phase
// functions (microwakeword_b200/csrc/mww_nn_dev.cuh), barriers modelled as phase boundaries.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../microwakeword_b200/csrc/mww_nn_mma.cuh"

using namespace mww;


// ---- host model of mma.sync.m16n8k8.tf32 for one warp (fragment layouts from the PTX ISA) -------------
namespace {
float tf32_of(uint32_t bits) { bits &= 0xFFFFE000u; float f; memcpy(&f, &bits, 4); return f; }   // hardware ignores the low 13 bits

void warp_mma_tf32(float (*c)[4], const uint32_t (*a)[4], const uint32_t (*b)[2]) {
    float A[16][8], B[8][8], C[16][8];
    for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, tig = lane & 3;
        A[g][tig] = tf32_of(a[lane][0]); A[g + 8][tig] = tf32_of(a[lane][1]);
        A[g][tig + 4] = tf32_of(a[lane][2]); A[g + 8][tig + 4] = tf32_of(a[lane][3]);
        B[tig][g] = tf32_of(b[lane][0]); B[tig + 4][g] = tf32_of(b[lane][1]);
        C[g][2 * tig] = c[lane][0]; C[g][2 * tig + 1] = c[lane][1];
        C[g + 8][2 * tig] = c[lane][2]; C[g + 8][2 * tig + 1] = c[lane][3];
    }
    for (int m = 0; m < 16; ++m)
        for (int n = 0; n < 8; ++n) {
            float acc = C[m][n];
            for (int k = 0; k < 8; ++k) acc += A[m][k] * B[k][n];
            C[m][n] = acc;
        }
    for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, tig = lane & 3;
        c[lane][0] = C[g][2 * tig]; c[lane][1] = C[g][2 * tig + 1];
        c[lane][2] = C[g + 8][2 * tig]; c[lane][3] = C[g + 8][2 * tig + 1];
    }
}

void warp_mma_3xtf32(float (*c)[4], const FragA *a, const FragB *b) {
    uint32_t ah[32][4], al[32][4], bh[32][2], bl[32][2];
    for (int l = 0; l < 32; ++l) {
        for (int i = 0; i < 4; ++i) { ah[l][i] = a[l].hi[i]; al[l][i] = a[l].lo[i]; }
        for (int i = 0; i < 2; ++i) { bh[l][i] = b[l].hi[i]; bl[l][i] = b[l].lo[i]; }
    }
    warp_mma_tf32(c, al, bh);
    warp_mma_tf32(c, ah, bl);
    warp_mma_tf32(c, ah, bh);
}

template <int L>
void emul_pointwise_mma(float *sm, const NnWeightsF32 &W) {
    constexpr int cin = kGeom[L].cin;
    for (int warp = 0; warp < kNnThreads / 32; ++warp) {
        const int t0 = 16 * pw_m_tile(warp), nt0 = pw_n_first(warp), ntc = pw_n_count(warp);
        const float *d = sm + kXFloats;
        const float *wsm = nn_pw_weight_buffer<L>(sm);
        float c[3][32][4] = {};
        for (int ks = 0; ks < cin / 8; ++ks) {
            FragA a[32];
            for (int lane = 0; lane < 32; ++lane) load_frag_a(d, kDLd, 8 * ks, t0, lane, a[lane]);
            for (int i = 0; i < ntc; ++i) {
                FragB b[32];
                for (int lane = 0; lane < 32; ++lane) load_frag_b(wsm, kWLd, 8 * ks, 8 * (nt0 + i), lane, b[lane]);
                warp_mma_3xtf32(c[i], a, b);
            }
        }
        for (int i = 0; i < ntc; ++i)
            for (int lane = 0; lane < 32; ++lane) pw_store_tile<L>(sm, W, t0, 8 * (nt0 + i), lane, c[i][lane]);
    }
}

void emul_first_conv_mma(float *sm, const NnWeightsF32 &W) {
    const float *feat = sm + kXFloats + kDFloats;
    static float c[kNnThreads / 32][4][32][4];
    memset(c, 0, sizeof c);
    for (int warp = 0; warp < kNnThreads / 32; ++warp) {        // part a
        const int t0 = 16 * fc_m_tile(warp);
        for (int ks = fc_k_begin(warp); ks < fc_k_end(warp); ++ks) {
            FragA a[32];
            for (int lane = 0; lane < 32; ++lane) fc_load_frag_a(feat, ks, t0, lane, a[lane]);
            for (int i = 0; i < 4; ++i) {
                FragB b[32];
                for (int lane = 0; lane < 32; ++lane) fc_load_frag_b(W.w0, ks, 8 * i, lane, b[lane]);
                warp_mma_3xtf32(c[warp][i], a, b);
            }
        }
        if (warp % 3 > 0)
            for (int i = 0; i < 4; ++i)
                for (int lane = 0; lane < 32; ++lane) fc_store_partial(sm, warp % 3, t0, 8 * i, lane, c[warp][i][lane]);
    }
    for (int warp = 0; warp < kNnThreads / 32; ++warp) {        // part b, after the barrier
        if (warp % 3 != 0) continue;
        for (int i = 0; i < 4; ++i)
            for (int lane = 0; lane < 32; ++lane) fc_finish_tile(sm, 16 * fc_m_tile(warp), 8 * i, lane, c[warp][i][lane]);
    }
}
}  // namespace

extern "C" int emul_nn_f32(const float *const *wp /* w0, dw_w[4], dw_b[4], pw_w[4], pw_b[4], head_w, head_b */,
                           float *state, float *pend, int n_pend, const void *rows, int n_rows, int rows_are_f32,
                           int n_streams, float *probs, int max_probs, float *logits) {
    NnWeightsF32 W;
    W.w0 = wp[0];
    for (int i = 0; i < 4; ++i) { W.dw_w[i] = wp[1 + i]; W.dw_b[i] = wp[5 + i]; W.pw_w[i] = wp[9 + i]; W.pw_b[i] = wp[13 + i]; }
    W.head_w = wp[17];
    W.head_b = wp[18];
    const int n_virtual = n_pend + n_rows;
    const int n_steps = n_virtual / 3;
    const size_t row_bytes = (size_t)kNumChannels * (rows_are_f32 ? 4 : 2);
    std::vector<float> smv(kNnSmemFloats);
    float *sm = smv.data();
    for (int s = 0; s < n_streams; ++s) {
        for (auto &v : smv) v = -1234.5f;   // poison
        float *my_state = state + (size_t)s * kStateFloats;
        float *my_pend = pend + (size_t)s * 2 * kNumChannels;
        NnInput in;
        in.ring0 = my_state; in.pend = my_pend; in.n_pend = n_pend;
        in.rows = static_cast<const char *>(rows) + (size_t)s * n_rows * row_bytes;
        in.n_rows = n_rows; in.rows_are_f32 = rows_are_f32;
#define ALL(stmt) for (int tid = 0; tid < kNnThreads; ++tid) { stmt; }
        ALL(nn_load_state(tid, sm, my_state));
        for (int step0 = 0; step0 < n_steps; step0 += kTT) {
            const int n = n_steps - step0 < kTT ? n_steps - step0 : kTT;
            ALL(nn_load_features(tid, sm, in, step0, n));
            emul_first_conv_mma(sm, W);
            ALL(nn_stage_pw_weights<0>(tid, sm, W)); ALL(nn_stage_pw_weights<1>(tid, sm, W));
            ALL(nn_depthwise<0>(tid, sm, W)); emul_pointwise_mma<0>(sm, W);
            ALL(nn_stage_pw_weights<2>(tid, sm, W));
            ALL(nn_depthwise<1>(tid, sm, W)); emul_pointwise_mma<1>(sm, W);
            ALL(nn_stage_pw_weights<3>(tid, sm, W));
            ALL(nn_depthwise<2>(tid, sm, W)); emul_pointwise_mma<2>(sm, W);
            ALL(nn_depthwise<3>(tid, sm, W)); emul_pointwise_mma<3>(sm, W);
            ALL(nn_head_partial(tid, sm, W));
            ALL(nn_head_finish(tid, sm, W, n, probs + (size_t)s * max_probs + step0, logits ? logits + (size_t)s * max_probs + step0 : nullptr));
            std::vector<float> tmp((size_t)kNnThreads * 5 * kShiftPerThread);
            ALL(nn_shift_read(tid, sm, n, *reinterpret_cast<float(*)[5][kShiftPerThread]>(&tmp[(size_t)tid * 5 * kShiftPerThread])));
            ALL(nn_shift_write(tid, sm, *reinterpret_cast<float(*)[5][kShiftPerThread]>(&tmp[(size_t)tid * 5 * kShiftPerThread])));
        }
        std::vector<NnTail> tails(kNnThreads);
        ALL(nn_tail_read(tid, in, n_steps, n_virtual, tails[tid]));
        ALL(nn_tail_write(tid, sm, my_state, my_pend, tails[tid]));
#undef ALL
    }
    return n_steps;
}

// ---- int8 path --------------------------------------------------------------------------------
#include "../../microwakeword_b200/csrc/mww_nn_i8_dev.cuh"

extern "C" int emul_nn_i8(const void *const *wp /* see order below */, const int32_t *zp12, const int32_t *head3, float in_scale,
                          int8_t *state, int8_t *pend, int n_pend, const void *rows, int n_rows, int row_type,
                          int n_streams, float *probs, int max_probs) {
    NnWeightsI8 W;
    int k = 0;
    W.w0 = (const int8_t *)wp[k++]; W.b0 = (const int32_t *)wp[k++]; W.m0 = (const int32_t *)wp[k++]; W.s0 = (const int32_t *)wp[k++];
    for (int i = 0; i < 4; ++i) {
        W.dw_w[i] = (const int8_t *)wp[k++]; W.dw_b[i] = (const int32_t *)wp[k++]; W.dw_m[i] = (const int32_t *)wp[k++]; W.dw_s[i] = (const int32_t *)wp[k++];
        W.pw_w[i] = (const int8_t *)wp[k++]; W.pw_b[i] = (const int32_t *)wp[k++]; W.pw_m[i] = (const int32_t *)wp[k++]; W.pw_s[i] = (const int32_t *)wp[k++];
    }
    W.head_w = (const int8_t *)wp[k++]; W.lut = (const int8_t *)wp[k++];
    W.head_bias = head3[0]; W.head_mult = head3[1]; W.head_shift = head3[2];
    memcpy(W.zp, zp12, sizeof W.zp);
    W.in_scale = in_scale;
    const int n_virtual = n_pend + n_rows;
    const int n_steps = n_virtual / 3;
    const size_t row_bytes = (size_t)kNumChannels * (row_type == 1 ? 4 : (row_type == 0 ? 2 : 1));
    std::vector<int32_t> smv(kNnSmemFloats);
    int32_t *sm = smv.data();
    for (int s = 0; s < n_streams; ++s) {
        for (auto &v : smv) v = 0x5A5A5A5A;
        int8_t *my_state = state + (size_t)s * kStateFloats;
        int8_t *my_pend = pend + (size_t)s * 2 * kNumChannels;
        NnInputI8 in;
        in.ring0 = my_state; in.pend = my_pend; in.n_pend = n_pend;
        in.rows = static_cast<const char *>(rows) + (size_t)s * n_rows * row_bytes;
        in.n_rows = n_rows; in.row_type = row_type;
#define ALL(stmt) for (int tid = 0; tid < kNnThreads; ++tid) { stmt; }
        ALL(nnq_load_state(tid, sm, my_state, W));
        for (int step0 = 0; step0 < n_steps; step0 += kTT) {
            const int n = n_steps - step0 < kTT ? n_steps - step0 : kTT;
            ALL(nnq_load_features(tid, sm, in, W, step0, n));
            std::vector<int32_t> fc((size_t)kNnThreads * 8);
            ALL(nnq_first_conv_a(tid, sm, W, *reinterpret_cast<int32_t(*)[2][4]>(&fc[(size_t)tid * 8])));
            ALL(nnq_first_conv_b(tid, sm, W, *reinterpret_cast<int32_t(*)[2][4]>(&fc[(size_t)tid * 8])));
            ALL(nnq_stage_pw_weights<0>(tid, sm, W)); ALL(nnq_depthwise<0>(tid, sm, W)); ALL(nnq_pointwise<0>(tid, sm, W));
            ALL(nnq_stage_pw_weights<1>(tid, sm, W)); ALL(nnq_depthwise<1>(tid, sm, W)); ALL(nnq_pointwise<1>(tid, sm, W));
            ALL(nnq_stage_pw_weights<2>(tid, sm, W)); ALL(nnq_depthwise<2>(tid, sm, W)); ALL(nnq_pointwise<2>(tid, sm, W));
            ALL(nnq_stage_pw_weights<3>(tid, sm, W)); ALL(nnq_depthwise<3>(tid, sm, W)); ALL(nnq_pointwise<3>(tid, sm, W));
            ALL(nnq_head_partial(tid, sm, W));
            ALL(nnq_head_finish(tid, sm, W, n, probs + (size_t)s * max_probs + step0));
            std::vector<float> tmp((size_t)kNnThreads * 5 * kShiftPerThread);
            ALL(nn_shift_read(tid, reinterpret_cast<const float *>(sm), n, *reinterpret_cast<float(*)[5][kShiftPerThread]>(&tmp[(size_t)tid * 5 * kShiftPerThread])));
            ALL(nn_shift_write(tid, reinterpret_cast<float *>(sm), *reinterpret_cast<float(*)[5][kShiftPerThread]>(&tmp[(size_t)tid * 5 * kShiftPerThread])));
        }
        std::vector<NnTailI8> tails(kNnThreads);
        ALL(nnq_tail_read(tid, in, W, n_steps, n_virtual, tails[tid]));
        ALL(nnq_tail_write(tid, sm, my_state, my_pend, W, tails[tid]));
#undef ALL
    }
    return n_steps;
}

extern "C" void emul_fill_state_i8(const int32_t *zp12, int8_t *state, int8_t *pend, int n_streams) {
    NnWeightsI8 W;
    memcpy(W.zp, zp12, sizeof W.zp);
    for (int s = 0; s < n_streams; ++s) {
        for (int e = 0; e < kStateFloats; ++e) state[(size_t)s * kStateFloats + e] = nnq_reset_value(W, e);
        for (int e = 0; e < 2 * kNumChannels; ++e) pend[(size_t)s * 2 * kNumChannels + e] = (int8_t)W.zp[0];
    }
}

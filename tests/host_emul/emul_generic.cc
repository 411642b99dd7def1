// tests/host_emul/emul_generic.cc -- TEST INFRASTRUCTURE: CPU execution of the run-time-geometry MixedNet phase functions
// (microwakeword_b200/csrc/mww_nn_generic.cuh) in exactly the order nn_generic_*_kernel issues them, every barrier a
// phase boundary, shared memory poisoned before each stream.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../microwakeword_b200/csrc/mww_nn_generic.cuh"

using namespace mww;

// Threads of a phase run one after the other; the order must not matter (a phase that reads what another thread of the
// SAME phase writes would be a race on the GPU).  tests/test_generic_arch.py runs every comparison in both orders.
static int g_reverse = 0;
extern "C" void emul_gen_thread_order(int reversed) { g_reverse = reversed; }
#define GEN_ALL(stmt) for (int t_ = 0; t_ < kGenThreads; ++t_) { const int tid = g_reverse ? kGenThreads - 1 - t_ : t_; stmt; }

namespace {
void advance(int &pos, int slots) { pos = pos + 1 == slots ? 0 : pos + 1; }
}  // namespace

// geometry numbers for the tests: {rc, state_elems, pend_cap, stride, sm_elems, macs_per_step, ring0, c_last}
extern "C" int emul_gen_arch(const int32_t *arch, int n_arch, long long *out8) {
    GenArch A;
    const int rc = gen_arch_from_tensor(arch, n_arch, &A);
    out8[0] = rc;
    if (rc) return rc;
    out8[1] = A.state_elems; out8[2] = A.pend_cap; out8[3] = A.stride; out8[4] = A.sm_elems; out8[5] = A.macs_per_step;
    out8[6] = A.ring0; out8[7] = A.c_last;
    return 0;
}

extern "C" int emul_gen_f32(const int32_t *arch, int n_arch, const float *const *wp /* w0, per block {dw_w, dw_b, pw_w, pw_b}, head_w, head_b */,
                            float *state, float *pend, int n_pend, const void *rows, int n_rows, int row_type, int n_streams,
                            float *probs, int max_probs) {
    GenArch A;
    if (gen_arch_from_tensor(arch, n_arch, &A)) return -1;
    GenWeightsF32 W = GenWeightsF32();
    int k = 0;
    W.w0 = wp[k++];
    for (int b = 0; b < A.n_blocks; ++b) { W.dw_w[b] = wp[k++]; W.dw_b[b] = wp[k++]; W.pw_w[b] = wp[k++]; W.pw_b[b] = wp[k++]; }
    W.head_w = wp[k++]; W.head_b = wp[k++];
    const int n_steps = (n_pend + n_rows) / A.stride;
    if (n_steps > max_probs) return -2;
    const size_t row_bytes = (size_t)kNumChannels * (row_type == 1 ? 4 : 2);
    std::vector<float> smv(A.sm_elems);
    float *sm = smv.data();
    for (int s = 0; s < n_streams; ++s) {
        for (auto &v : smv) v = -1234.5f;   // poison
        float *my_state = state + (size_t)s * A.state_elems;
        float *my_pend = pend + (size_t)s * A.pend_cap * kNumChannels;
        GenInput<float> in;
        in.state = my_state; in.pend = my_pend; in.n_pend = n_pend;
        in.rows = static_cast<const char *>(rows) + (size_t)s * n_rows * row_bytes;
        in.n_rows = n_rows; in.row_type = row_type;
        int pos[kGenMaxBlocks + 1] = {0};
        GEN_ALL(gen_f32_load_state(tid, sm, A, my_state));
        for (int t = 0; t < n_steps; ++t) {
            GEN_ALL(gen_f32_window(tid, sm, A, in, t));
            GEN_ALL(gen_f32_first_conv(tid, sm, A, W));
            for (int b = 0; b < A.n_blocks; ++b) {
                GEN_ALL(gen_f32_depthwise(tid, sm, A, W, b, pos[b]));
                advance(pos[b], A.kmax[b]);
                GEN_ALL(gen_f32_pointwise(tid, sm, A, W, b));
            }
            GEN_ALL(gen_f32_head_partial(tid, sm, A, W, pos[A.n_blocks]));
            advance(pos[A.n_blocks], A.head_rows);
            GEN_ALL(gen_f32_head_finish(tid, sm, A, W, probs + (size_t)s * max_probs + t));
        }
        GEN_ALL(gen_f32_tail_gather(tid, sm, A, in, n_steps));
        GEN_ALL(gen_f32_tail_store(tid, sm, A, my_state, my_pend, A.ring0 + n_pend + n_rows - A.stride * n_steps, pos));
    }
    return n_steps;
}

extern "C" int emul_gen_i8(const int32_t *arch, int n_arch, const void *const *wp /* w0,b0,m0,s0, per block {dw w,b,m,s, pw w,b,m,s}, head_w, lut */,
                           const int32_t *zps, const int32_t *head3, float in_scale, int8_t *state, int8_t *pend, int n_pend,
                           const void *rows, int n_rows, int row_type, int n_streams, float *probs, int max_probs) {
    GenArch A;
    if (gen_arch_from_tensor(arch, n_arch, &A)) return -1;
    GenWeightsI8 W = GenWeightsI8();
    int k = 0;
    W.w0 = (const int8_t *)wp[k++]; W.b0 = (const int32_t *)wp[k++]; W.m0 = (const int32_t *)wp[k++]; W.s0 = (const int32_t *)wp[k++];
    for (int b = 0; b < A.n_blocks; ++b) {
        W.dw_w[b] = (const int8_t *)wp[k++]; W.dw_b[b] = (const int32_t *)wp[k++]; W.dw_m[b] = (const int32_t *)wp[k++]; W.dw_s[b] = (const int32_t *)wp[k++];
        W.pw_w[b] = (const int8_t *)wp[k++]; W.pw_b[b] = (const int32_t *)wp[k++]; W.pw_m[b] = (const int32_t *)wp[k++]; W.pw_s[b] = (const int32_t *)wp[k++];
    }
    W.head_w = (const int8_t *)wp[k++]; W.lut = (const int8_t *)wp[k++];
    W.head_bias = head3[0]; W.head_mult = head3[1]; W.head_shift = head3[2];
    memcpy(W.zp, zps, sizeof(int32_t) * (4 + 2 * A.n_blocks));
    W.in_scale = in_scale;
    const int n_steps = (n_pend + n_rows) / A.stride;
    if (n_steps > max_probs) return -2;
    const size_t row_bytes = (size_t)kNumChannels * (row_type == 1 ? 4 : (row_type == 0 ? 2 : 1));
    std::vector<int32_t> smv(A.sm_elems);
    int32_t *sm = smv.data();
    for (int s = 0; s < n_streams; ++s) {
        for (auto &v : smv) v = 0x5A5A5A5A;
        int8_t *my_state = state + (size_t)s * A.state_elems;
        int8_t *my_pend = pend + (size_t)s * A.pend_cap * kNumChannels;
        GenInput<int8_t> in;
        in.state = my_state; in.pend = my_pend; in.n_pend = n_pend;
        in.rows = static_cast<const char *>(rows) + (size_t)s * n_rows * row_bytes;
        in.n_rows = n_rows; in.row_type = row_type;
        int pos[kGenMaxBlocks + 1] = {0};
        GEN_ALL(gen_i8_load_state(tid, sm, A, my_state));
        for (int t = 0; t < n_steps; ++t) {
            GEN_ALL(gen_i8_window(tid, sm, A, W, in, t));
            GEN_ALL(gen_i8_first_conv(tid, sm, A, W));
            for (int b = 0; b < A.n_blocks; ++b) {
                GEN_ALL(gen_i8_depthwise(tid, sm, A, W, b, pos[b]));
                advance(pos[b], A.kmax[b]);
                GEN_ALL(gen_i8_pointwise(tid, sm, A, W, b));
            }
            GEN_ALL(gen_i8_head_partial(tid, sm, A, W, pos[A.n_blocks]));
            advance(pos[A.n_blocks], A.head_rows);
            GEN_ALL(gen_i8_head_finish(tid, sm, A, W, probs + (size_t)s * max_probs + t));
        }
        GEN_ALL(gen_i8_tail_gather(tid, sm, A, W, in, n_steps));
        GEN_ALL(gen_i8_tail_store(tid, sm, A, my_state, my_pend, A.ring0 + n_pend + n_rows - A.stride * n_steps, pos));
    }
    return n_steps;
}

extern "C" int emul_gen_fill_state_i8(const int32_t *arch, int n_arch, const int32_t *zps, int8_t *state, int8_t *pend, int n_streams) {
    GenArch A;
    if (gen_arch_from_tensor(arch, n_arch, &A)) return -1;
    GenWeightsI8 W = GenWeightsI8();
    memcpy(W.zp, zps, sizeof(int32_t) * (4 + 2 * A.n_blocks));
    for (int s = 0; s < n_streams; ++s) {
        for (int e = 0; e < A.state_elems; ++e) state[(size_t)s * A.state_elems + e] = gen_i8_reset_value(A, W, e);
        for (int e = 0; e < A.pend_cap * kNumChannels; ++e) pend[(size_t)s * A.pend_cap * kNumChannels + e] = (int8_t)W.zp[0];
    }
    return 0;
}

// mww_nn_live.cu -- sm_100a kernel + launcher of the live-step MixedNet (one model step for many streams per
// launch; phase functions and rationale: mww_nn_live.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "mww_kernels.h"
#include "mww_nn_live.cuh"

namespace mww {

// persistent CTAs: each walks groups of 32 streams; 1x1 weights are staged once per CTA
__global__ void __launch_bounds__(kLiveThreads, 2)
nn_f32_live_kernel(NnWeightsF32 W, float *__restrict__ state, float *__restrict__ pend, int n_pend, const void *__restrict__ rows,
                   long long rows_stream_stride_bytes, int rows_are_f32, float *__restrict__ probs, long long probs_stride,
                   int n_streams, LiveHeads heads) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    LiveInput in;
    in.state = state; in.pend = pend; in.n_pend = n_pend; in.rows = rows;
    in.rows_stream_stride_bytes = rows_stream_stride_bytes; in.rows_are_f32 = rows_are_f32;
    live_load_weights(tid, sm, W);
    __syncthreads();
    const int n_groups = (n_streams + kLiveStreams - 1) / kLiveStreams;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long s0 = (long long)g * kLiveStreams;
        const int n_valid = min(kLiveStreams, n_streams - (int)s0);
        float keep[4][kLiveKeep];
        // every ring is requested from L2 two to three stages before its register loads (mww_nn_live.cuh, "L2 prefetch");
        // at most ~9.7 KB per stream are requested-but-unread at any point of the chain
        live_prefetch_rings<1, 2>(tid, state, s0, n_valid);
        live_build_a(tid, sm, in, s0, n_valid, keep);
        __syncthreads();
        live_prefetch_rings<3, 3>(tid, state, s0, n_valid);
        live_write_tail(tid, state, pend, s0, n_valid, keep);
        live_first_conv_mma(tid, sm, W.w0, 32);
        __syncthreads();
        live_depthwise<0>(tid, sm, W, state, s0, n_valid, heads.h[0]); __syncthreads();
        live_pointwise_mma<0>(tid, sm, W); __syncthreads();
        live_depthwise<1>(tid, sm, W, state, s0, n_valid, heads.h[1]); __syncthreads();
        live_prefetch_rings<4, 4>(tid, state, s0, n_valid);
        live_pointwise_mma<1>(tid, sm, W); __syncthreads();
        live_depthwise<2>(tid, sm, W, state, s0, n_valid, heads.h[2]); __syncthreads();
        live_prefetch_rings<5, 5>(tid, state, s0, n_valid);
        live_pointwise_mma<2>(tid, sm, W); __syncthreads();
        live_depthwise<3>(tid, sm, W, state, s0, n_valid, heads.h[3]); __syncthreads();
        live_prefetch_next_window(tid, state, pend, rows, rows_stream_stride_bytes, rows_are_f32 ? 480u : 240u,
                                  (long long)(g + gridDim.x) * kLiveStreams, n_streams);
        live_pointwise_mma<3>(tid, sm, W); __syncthreads();
        live_head_partial(tid, sm, W, state, s0, n_valid, heads.h[4]);
        __syncthreads();
        live_head_finish(tid, sm, W, s0, n_valid, probs, probs_stride);
        // the next group's D / H writes are separated from these reads by the barriers at its top
    }
}

// v2 (mww_nn_live.cuh "warp-specialised live step"): threads 0..255 = layer chain, 256..767 = ring streamers
constexpr int kBarChain = 1, kBarFull0 = 2, kBarEmpty0 = 4, kBarAEmpty = 6;   // named barriers: full / empty come in pairs (buffer 0, 1)
__global__ void __launch_bounds__(kLive2Threads, 1)
nn_f32_live2_kernel(NnWeightsF32 W, float *__restrict__ state, float *__restrict__ pend, int n_pend, const void *__restrict__ rows,
                    long long rows_stream_stride_bytes, int rows_are_f32, float *__restrict__ probs, long long probs_stride,
                    int n_streams, LiveHeads heads, int debug_mode) {
    // debug_mode (MWW_LIVE_MODE, timing experiments only -- results are garbage): 1 = streamers skip their loads, 2 = the chain
    // skips its work; both keep the barrier protocol, so the other side runs at its own pace
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    for (int L = 1; L < 4; ++L) {                                   // 1x1 weights of blocks 1..3, resident for the whole launch
        float *dst = sm + (L == 1 ? kLiveOffPw1 : (L == 2 ? kLiveOffPw2 : kLiveOffPw3));
        for (int e = tid; e < 64 * 64; e += kLive2Threads) dst[(e >> 6) * kWLd + (e & 63)] = W.pw_w[L][e];
    }
    live2_stage_chain_tables(tid, kLive2Threads, sm, W);
    __syncthreads();
    const int n_groups = (n_streams + kLiveStreams - 1) / kLiveStreams;
    LiveInput in;
    in.state = state; in.pend = pend; in.n_pend = n_pend; in.rows = rows;
    in.rows_stream_stride_bytes = rows_stream_stride_bytes; in.rows_are_f32 = rows_are_f32;
    if (tid >= kLive2ChainThreads) {
        // ---- streamers: P of group k into buffer k & 1 (at most two groups ahead of the chain), then the group's first-conv
        // window into A as soon as the chain's first conv of the previous group has let go of it ----
        const int st = tid - kLive2ChainThreads;
        int k = 0;
        for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++k) {
            const int buf = k & 1;
            if (k >= 2) bar_sync(kBarEmpty0 + buf, kLive2Threads);                  // the chain is done with this P buffer
            const long long s0 = (long long)g * kLiveStreams;
            const int n_valid = min(kLiveStreams, n_streams - (int)s0);
            if (debug_mode != 1) live2_stream_group(st, W, state, s0, n_valid, heads, sm + kLive2OffP + buf * kLive2PFloats);
            if (k >= 1) bar_sync(kBarAEmpty, kLive2Threads);                        // first conv + tail of group k - 1 have read A
            if (debug_mode != 1) live2_build_a<kLive2StreamThreads>(st, sm, in, s0, n_valid);
            bar_arrive(kBarFull0 + buf, kLive2Threads);
        }
        return;
    }
    // ---- chain: shared memory in, ring rows / scores out; no global load on its critical path ----
    int k = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++k) {
        const int buf = k & 1;
        const float *p_buf = sm + kLive2OffP + buf * kLive2PFloats;
        const long long s0 = (long long)g * kLiveStreams;
        const int n_valid = min(kLiveStreams, n_streams - (int)s0);
        bar_sync(kBarFull0 + buf, kLive2Threads);                                   // P and A of this group are complete
        if (debug_mode == 2) {
            bar_arrive(kBarAEmpty, kLive2Threads);
            bar_arrive(kBarEmpty0 + buf, kLive2Threads);
            continue;
        }
        live_first_conv_mma(tid, sm, sm + kLive2OffW0, kLive2W0Pitch);
        bar_sync(kBarChain, kLive2ChainThreads);                                    // H complete, every warp is done reading A as an operand
        live2_write_tail(tid, sm, in, state, pend, s0, n_valid);                    // new first-conv ring + pending rows, from A
        bar_arrive(kBarAEmpty, kLive2Threads);
        live2_dw_from_p<0>(tid, sm, state, s0, n_valid, heads.h[0], p_buf); bar_sync(kBarChain, kLive2ChainThreads);
        live_pointwise_mma<0, true>(tid, sm, W); bar_sync(kBarChain, kLive2ChainThreads);
        live2_dw_from_p<1>(tid, sm, state, s0, n_valid, heads.h[1], p_buf); bar_sync(kBarChain, kLive2ChainThreads);
        live_pointwise_mma<1, true>(tid, sm, W); bar_sync(kBarChain, kLive2ChainThreads);
        live2_dw_from_p<2>(tid, sm, state, s0, n_valid, heads.h[2], p_buf); bar_sync(kBarChain, kLive2ChainThreads);
        live_pointwise_mma<2, true>(tid, sm, W); bar_sync(kBarChain, kLive2ChainThreads);
        live2_dw_from_p<3>(tid, sm, state, s0, n_valid, heads.h[3], p_buf); bar_sync(kBarChain, kLive2ChainThreads);
        live_pointwise_mma<3, true>(tid, sm, W); bar_sync(kBarChain, kLive2ChainThreads);
        live2_dw_from_p<4>(tid, sm, state, s0, n_valid, heads.h[4], p_buf);
        bar_arrive(kBarEmpty0 + buf, kLive2Threads);                                // last read of this P buffer
        bar_sync(kBarChain, kLive2ChainThreads);
        live_head_finish_b(tid, sm, sm[kLive2OffSmall + kLive2SmallHeadBias], s0, n_valid, probs, probs_stride);
        bar_sync(kBarChain, kLive2ChainThreads);                                    // D is rewritten by the next group
    }
}

cudaError_t launch_nn_f32_live(const NnWeightsF32 &W, float *state, float *pend, int n_pend, const void *rows,
                               long long rows_stream_stride_bytes, int rows_are_f32, float *probs, long long probs_stride,
                               int n_streams, const LiveHeads &heads, int sm_count, int variant, cudaStream_t st) {
    if (n_streams <= 0) return cudaSuccess;
    static bool attr_done[64] = {};
    const bool v1 = variant != 2;       // 2 = the warp-specialised kernel (MWW_LIVE_V2 at mww_create): the r02 measurement instrument, same speed
    static const int debug_mode = getenv("MWW_LIVE_MODE") ? atoi(getenv("MWW_LIVE_MODE")) : 0;
    if (first_launch_on_this_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(nn_f32_live_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLiveSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nn_f32_live2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLive2SmemBytes);
        if (e != cudaSuccess) return e;
    }
    const int n_groups = (n_streams + kLiveStreams - 1) / kLiveStreams;
    if (v1) {
        const int grid = std::min(n_groups, 2 * sm_count);
        nn_f32_live_kernel<<<grid, kLiveThreads, kLiveSmemBytes, st>>>(W, state, pend, n_pend, rows, rows_stream_stride_bytes, rows_are_f32, probs,
                                                                       probs_stride, n_streams, heads);
    } else {
        const int grid = std::min(n_groups, sm_count);
        nn_f32_live2_kernel<<<grid, kLive2Threads, kLive2SmemBytes, st>>>(W, state, pend, n_pend, rows, rows_stream_stride_bytes, rows_are_f32, probs,
                                                                          probs_stride, n_streams, heads, debug_mode);
    }
    return cudaGetLastError();
}

__global__ void nn_live_canonicalise_kernel(float *__restrict__ state, int n_streams, LiveHeads heads) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_streams * 288) return;
    live_canonicalise_column(state, e / 288, (int)(e % 288), heads);
}

cudaError_t launch_nn_live_canonicalise(float *state, int n_streams, const LiveHeads &heads, cudaStream_t st) {
    if (n_streams <= 0) return cudaSuccess;
    const long long total = (long long)n_streams * 288;
    nn_live_canonicalise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(state, n_streams, heads);
    return cudaGetLastError();
}

}  // namespace mww

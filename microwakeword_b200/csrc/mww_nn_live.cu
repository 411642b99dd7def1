// mww_nn_live.cu -- sm_100a kernel + launcher of the live-step MixedNet (one model step for many streams per
// launch; phase functions and rationale: mww_nn_live.cuh).
#include <cuda_runtime.h>

#include <algorithm>

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
        live_first_conv_mma(tid, sm, W);
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

cudaError_t launch_nn_f32_live(const NnWeightsF32 &W, float *state, float *pend, int n_pend, const void *rows,
                               long long rows_stream_stride_bytes, int rows_are_f32, float *probs, long long probs_stride,
                               int n_streams, const LiveHeads &heads, int sm_count, cudaStream_t st) {
    if (n_streams <= 0) return cudaSuccess;
    static bool attr_done[64] = {};
    if (first_launch_on_this_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(nn_f32_live_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLiveSmemBytes);
        if (e != cudaSuccess) return e;
    }
    const int n_groups = (n_streams + kLiveStreams - 1) / kLiveStreams;
    const int grid = std::min(n_groups, 2 * sm_count);
    nn_f32_live_kernel<<<grid, kLiveThreads, kLiveSmemBytes, st>>>(W, state, pend, n_pend, rows, rows_stream_stride_bytes, rows_are_f32, probs,
                                                                   probs_stride, n_streams, heads);
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

// tools/probes/tcgen05_pointwise_probe.cu -- MEASUREMENT PROBE, not product code (VERDICT r01 next-5 iii).
//
// Question: would the clip kernel's 1x1 projections (3xTF32, K = 64, N = 64) be faster on tcgen05 (tcgen05.mma kind::tf32,
// operands in shared memory behind matrix descriptors, fp32 accumulators in TMEM) than on warp-level mma.sync, once the cost
// of putting the CUDA-core-produced A operand where the instruction wants it is counted?
//
// Both kernels do the same work per tile, in a loop, with one CTA per SM slot:
//   produce   A[128 x 64] = f(tile, row, k)  on CUDA cores (stands in for the depthwise output of 128 model steps)
//   contract  D[128 x 64] = A . B^T with the 3xTF32 split (lo*hi + hi*lo + hi*hi), B = fixed weights [64 x 64]
//   epilogue  bias + ReLU, 128 x 64 floats stored to global
//   * tcgen05 variant: A written ONCE as hi / lo planes in the canonical K-major no-swizzle core-matrix layout, fence to the
//     async proxy, one thread issues 24 tcgen05.mma (8 k-steps x 3 products) + tcgen05.commit -> mbarrier, all four warps
//     read their 32 accumulator lanes back with tcgen05.ld (32x32b.x64).
//   * mma.sync variant: A written as plain fp32 [k][t] (pitch 136), every warp splits its fragments on use (the r01 scheme).
// The probe checks both results against an fp64 reference, then times T tiles per CTA on a full grid (CUDA events).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_probe tcgen05_pointwise_probe.cu && ./tcgen05_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int M = 128, N = 64, K = 64;
constexpr int kThreads = 128;

__host__ __device__ inline float a_value(int tile, int row, int k) {
    // cheap, deterministic, order-1 magnitude with full mantissas
    const unsigned h = (unsigned)(tile * 7919 + row * 131 + k * 17) * 2654435761u;
    return (float)(h >> 8) * (1.0f / 16777216.0f) * 2.0f - 0.75f;
}

__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// ------------------------------------------------------------------------------------------------ tcgen05 variant
// canonical K-major, no swizzle: core matrix = 8 rows x 16 bytes (4 tf32), stored as 128 contiguous bytes;
// element (r, k) of a [rows x 64] tile lives at  (r / 8) * SBO + (k / 4) * LBO + (r % 8) * 16 + (k % 4) * 4
constexpr uint32_t kLBO = 128;                 // next core matrix along K
constexpr uint32_t kSBO = 16 * 128;            // next 8-row group: 16 core matrices (K = 64) further
__host__ __device__ inline uint32_t canon_off(int r, int k) { return (uint32_t)(r >> 3) * kSBO + (uint32_t)(k >> 2) * kLBO + (uint32_t)(r & 7) * 16 + (uint32_t)(k & 3) * 4; }

__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);                 // start address, bits [0, 14)
    d |= (uint64_t)(kLBO >> 4) << 16;                            // leading dimension byte offset, bits [16, 30)
    d |= (uint64_t)(kSBO >> 4) << 32;                            // stride dimension byte offset, bits [32, 46)
    d |= (uint64_t)1 << 46;                                      // descriptor version 1 (sm_100)
    return d;                                                    // base offset 0, layout type 0 = SWIZZLE_NONE
}
// instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
                 : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
probe_tcgen05(const float *__restrict__ w /* [N][K] */, const float *__restrict__ bias, float *__restrict__ out, int tiles_per_cta, int store_all) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *a_hi = reinterpret_cast<float *>(smem);                       // 32 KB each
    float *a_lo = a_hi + M * K;
    float *b_hi = a_lo + M * K;                                          // 16 KB each
    float *b_lo = b_hi + N * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    // weights once: hi / lo planes in the canonical layout
    for (int e = tid; e < N * K; e += kThreads) {
        const int n = e / K, k = e - n * K;
        float hi, lo;
        split_tf32(w[e], hi, lo);
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(b_hi) + canon_off(n, k)) = hi;
        *reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(b_lo) + canon_off(n, k)) = lo;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_smem)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;
    const uint32_t a_hi_s = (uint32_t)__cvta_generic_to_shared(a_hi), a_lo_s = (uint32_t)__cvta_generic_to_shared(a_lo);
    const uint32_t b_hi_s = (uint32_t)__cvta_generic_to_shared(b_hi), b_lo_s = (uint32_t)__cvta_generic_to_shared(b_lo);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t phase = 0;
    for (int it = 0; it < tiles_per_cta; ++it) {
        const int tile = blockIdx.x * tiles_per_cta + it;
        // produce: thread = row, all 64 k (the depthwise stage's "one lane per channel" has the same one-writer-per-element shape)
        {
            const int r = tid;
#pragma unroll 8
            for (int k = 0; k < K; k += 4) {
                float4 h, l;
                split_tf32(a_value(tile, r, k), h.x, l.x); split_tf32(a_value(tile, r, k + 1), h.y, l.y);
                split_tf32(a_value(tile, r, k + 2), h.z, l.z); split_tf32(a_value(tile, r, k + 3), h.w, l.w);
                *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(a_hi) + canon_off(r, k)) = h;     // one 16-byte core-matrix row
                *reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(a_lo) + canon_off(r, k)) = l;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");        // generic-proxy writes -> visible to the tensor core's async proxy
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < K / 8; ++ks) {                               // one MMA = K 8 = two core matrices along K
                const uint32_t adv = ks * 2 * kLBO;
                mma_tf32_ss(tmem_d, smem_desc(a_lo_s + adv), smem_desc(b_hi_s + adv), ks > 0);
                mma_tf32_ss(tmem_d, smem_desc(a_hi_s + adv), smem_desc(b_lo_s + adv), 1);
                mma_tf32_ss(tmem_d, smem_desc(a_hi_s + adv), smem_desc(b_hi_s + adv), 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar_s) : "memory");
        }
        // everyone waits for the accumulator
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(bar_s), "r"(phase) : "memory");
            }
            phase ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 (= rows); 64 columns in two loads of 32
        float *o = out + ((size_t)(store_all ? tile : blockIdx.x) * M + tid) * N;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                         "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                           "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                           "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 r4;
                r4.x = fmaxf(__uint_as_float(v[j]) + bias[half * 32 + j], 0.f); r4.y = fmaxf(__uint_as_float(v[j + 1]) + bias[half * 32 + j + 1], 0.f);
                r4.z = fmaxf(__uint_as_float(v[j + 2]) + bias[half * 32 + j + 2], 0.f); r4.w = fmaxf(__uint_as_float(v[j + 3]) + bias[half * 32 + j + 3], 0.f);
                *reinterpret_cast<float4 *>(o + half * 32 + j) = r4;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();                                    // accumulator and A planes are free for the next tile
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(64));
}

// ------------------------------------------------------------------------------------------------ tcgen05, SWIZZLE_128B
// The layout the product kernel uses: K is cut into slots of 32 tf32 = 128 bytes; a slot holds [rows][128 B] with the 16-byte
// chunks of a row XOR-ed by (row % 8) (the hardware's 128-byte swizzle), so a warp that stores one row -- 32 consecutive k for
// one time step, which is what a lane-per-channel producer does -- writes 128 contiguous bytes: no bank conflicts, unlike the
// 8-way conflicts the no-swizzle core-matrix layout gives that access pattern.  Descriptor: layout type 2, stride byte offset
// 1024 (8 rows), start address advanced by 32 bytes per K = 8 step inside the atom; slot base 1024-byte aligned.
__host__ __device__ inline uint32_t sw128_off(int r, int kk /* 0..31 inside the slot */) {
    return (uint32_t)r * 128 + (uint32_t)(((kk >> 2) ^ (r & 7)) << 4) + (uint32_t)(kk & 3) * 4;
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                                      // leading byte offset: unused for swizzled K-major (one atom wide)
    d |= (uint64_t)(1024 >> 4) << 32;                            // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                                      // descriptor version 1
    d |= (uint64_t)2 << 61;                                      // SWIZZLE_128B
    return d;
}
__global__ void __launch_bounds__(kThreads, 1)
probe_tcgen05_sw128(const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ out, int tiles_per_cta, int store_all) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // slot s of A: hi plane at s * 32 KB, lo plane at s * 32 KB + 16 KB;  slot s of B: 64 KB + s * 16 KB (+ 8 KB for lo)
    unsigned char *a_base = smem, *b_base = smem + 65536;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < N * K; e += kThreads) {
        const int n = e / K, k = e - n * K;
        float hi, lo;
        split_tf32(w[e], hi, lo);
        unsigned char *slot = b_base + (k >> 5) * 16384;
        *reinterpret_cast<float *>(slot + sw128_off(n, k & 31)) = hi;
        *reinterpret_cast<float *>(slot + 8192 + sw128_off(n, k & 31)) = lo;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_smem)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;
    const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(a_base), b_s = (uint32_t)__cvta_generic_to_shared(b_base);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t phase = 0;
    for (int it = 0; it < tiles_per_cta; ++it) {
        const int tile = blockIdx.x * tiles_per_cta + it;
        // produce with lanes along k (a warp stores whole 128-byte rows): thread -> (k = tid % 64, rows tid / 64 + 2 i)
        {
            const int k = tid & 63, r0 = tid >> 6;
            unsigned char *slot = a_base + (k >> 5) * 32768;
#pragma unroll 8
            for (int r = r0; r < M; r += 2) {
                float hi, lo;
                split_tf32(a_value(tile, r, k), hi, lo);
                *reinterpret_cast<float *>(slot + sw128_off(r, k & 31)) = hi;
                *reinterpret_cast<float *>(slot + 16384 + sw128_off(r, k & 31)) = lo;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint32_t a_hi = a_s + (ks >> 2) * 32768 + (ks & 3) * 32, a_lo = a_hi + 16384;
                const uint32_t b_hi = b_s + (ks >> 2) * 16384 + (ks & 3) * 32, b_lo = b_hi + 8192;
                mma_tf32_ss(tmem_d, smem_desc_sw128(a_lo), smem_desc_sw128(b_hi), ks > 0);
                mma_tf32_ss(tmem_d, smem_desc_sw128(a_hi), smem_desc_sw128(b_lo), 1);
                mma_tf32_ss(tmem_d, smem_desc_sw128(a_hi), smem_desc_sw128(b_hi), 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar_s) : "memory");
        }
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(bar_s), "r"(phase) : "memory");
            }
            phase ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        float *o = out + ((size_t)(store_all ? tile : blockIdx.x) * M + tid) * N;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                         "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                           "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                           "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 r4;
                r4.x = fmaxf(__uint_as_float(v[j]) + bias[half * 32 + j], 0.f); r4.y = fmaxf(__uint_as_float(v[j + 1]) + bias[half * 32 + j + 1], 0.f);
                r4.z = fmaxf(__uint_as_float(v[j + 2]) + bias[half * 32 + j + 2], 0.f); r4.w = fmaxf(__uint_as_float(v[j + 3]) + bias[half * 32 + j + 3], 0.f);
                *reinterpret_cast<float4 *>(o + half * 32 + j) = r4;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(64));
}

// ------------------------------------------------------------------------------------------------ mma.sync variant
constexpr int kPitchA = 136;      // [k][t] pitch: 136 = 8 (mod 32): conflict-free A fragments
constexpr int kPitchB = 72;       // [k][n] pitch
__device__ __forceinline__ void mma_1688(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void __launch_bounds__(kThreads, 1)
probe_mma_sync(const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ out, int tiles_per_cta, int store_all) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *a = reinterpret_cast<float *>(smem);               // [K][kPitchA]
    float *b = a + K * kPitchA;                               // [K][kPitchB]  (b[k][n] = w[n][k])
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
    for (int e = tid; e < N * K; e += kThreads) { const int n = e / K, k = e - n * K; b[k * kPitchB + n] = w[e]; }
    __syncthreads();
    for (int it = 0; it < tiles_per_cta; ++it) {
        const int tile = blockIdx.x * tiles_per_cta + it;
        {
            const int r = tid;
#pragma unroll 8
            for (int k = 0; k < K; ++k) a[k * kPitchA + r] = a_value(tile, r, k);
        }
        __syncthreads();
        // warp = 32 rows (two m-tiles) x 64 columns (eight n-tiles)
        float *o = out + ((size_t)(store_all ? tile : blockIdx.x) * M) * N;
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
            const int t0 = warp * 32 + mt * 16;
            float c[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
#pragma unroll 2
            for (int ks = 0; ks < K / 8; ++ks) {
                uint32_t ah[4], al[4];
                const float *p = a + (8 * ks + tig) * kPitchA + t0 + g;
                float h, l;
                split_tf32(p[0], h, l); ah[0] = __float_as_uint(h); al[0] = __float_as_uint(l);
                split_tf32(p[8], h, l); ah[1] = __float_as_uint(h); al[1] = __float_as_uint(l);
                split_tf32(p[4 * kPitchA], h, l); ah[2] = __float_as_uint(h); al[2] = __float_as_uint(l);
                split_tf32(p[4 * kPitchA + 8], h, l); ah[3] = __float_as_uint(h); al[3] = __float_as_uint(l);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    uint32_t bh[2], bl[2];
                    const float *q = b + (8 * ks + tig) * kPitchB + 8 * nt + g;
                    split_tf32(q[0], h, l); bh[0] = __float_as_uint(h); bl[0] = __float_as_uint(l);
                    split_tf32(q[4 * kPitchB], h, l); bh[1] = __float_as_uint(h); bl[1] = __float_as_uint(l);
                    mma_1688(c[nt], al, bh); mma_1688(c[nt], ah, bl); mma_1688(c[nt], ah, bh);
                }
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = t0 + g + ((i & 2) ? 8 : 0), col = 8 * nt + 2 * tig + (i & 1);
                    o[(size_t)row * N + col] = fmaxf(c[nt][i] + bias[col], 0.f);
                }
        }
        __syncthreads();
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<float> w(N * K), bias(N);
    for (int i = 0; i < N * K; ++i) w[i] = (float)((i * 2654435761u) >> 8) * (1.0f / 16777216.0f) - 0.5f;
    for (int i = 0; i < N; ++i) bias[i] = 0.01f * (i - 20);
    float *d_w, *d_b, *d_out;
    const int check_tiles = 2 * sms;
    CK(cudaMalloc(&d_w, w.size() * 4)); CK(cudaMalloc(&d_b, bias.size() * 4)); CK(cudaMalloc(&d_out, (size_t)check_tiles * M * N * 4));
    CK(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_b, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    const int smem_tc = (2 * M * K + 2 * N * K) * 4 + 1024, smem_mma = (K * kPitchA + K * kPitchB) * 4;
    CK(cudaFuncSetAttribute(probe_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc));
    CK(cudaFuncSetAttribute(probe_mma_sync, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_mma));
    CK(cudaFuncSetAttribute(probe_tcgen05_sw128, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc));
    const char *names[3] = {"tcgen05 kind::tf32 (no swizzle)", "mma.sync m16n8k8", "tcgen05 kind::tf32 (SWIZZLE_128B, K slots of 32)"};
    // ---- correctness: 2 tiles per CTA, every tile stored, against an fp64 reference
    std::vector<float> got((size_t)check_tiles * M * N);
    for (int variant = 0; variant < 3; ++variant) {
        CK(cudaMemset(d_out, 0, got.size() * 4));
        if (variant == 0) probe_tcgen05<<<sms, kThreads, smem_tc>>>(d_w, d_b, d_out, 2, 1);
        else if (variant == 1) probe_mma_sync<<<sms, kThreads, smem_mma>>>(d_w, d_b, d_out, 2, 1);
        else probe_tcgen05_sw128<<<sms, kThreads, smem_tc>>>(d_w, d_b, d_out, 2, 1);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int tile = 0; tile < check_tiles; tile += (tile < 4 ? 1 : 37))
            for (int r = 0; r < M; ++r)
                for (int n = 0; n < N; ++n) {
                    double acc = bias[n];
                    for (int k = 0; k < K; ++k) acc += (double)a_value(tile, r, k) * (double)w[n * K + k];
                    const double want = acc > 0 ? acc : 0;
                    worst = fmax(worst, fabs(want - (double)got[((size_t)tile * M + r) * N + n]));
                }
        printf("%s: max |error| vs fp64 reference %.3g (3xTF32 bound ~1e-5)\n", names[variant], worst);
    }
    // ---- timing: T tiles per CTA, only the last tile of a CTA is kept (stores stay in L2), CUDA events
    const int T = 2000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int variant = 0; variant < 3; ++variant) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            if (variant == 0) probe_tcgen05<<<sms, kThreads, smem_tc>>>(d_w, d_b, d_out, T, 0);
            else if (variant == 1) probe_mma_sync<<<sms, kThreads, smem_mma>>>(d_w, d_b, d_out, T, 0);
            else probe_tcgen05_sw128<<<sms, kThreads, smem_tc>>>(d_w, d_b, d_out, T, 0);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us_per_tile = ms * 1e3 / T;
        printf("%s: %.3f us per 128x64x64 3xTF32 tile per CTA (1 CTA/SM, %d SMs, produce + contract + epilogue); %.1f TFLOP/s chip-wide (3 x 2MNK)\n",
               names[variant], us_per_tile, sms, 3.0 * 2 * M * N * K * sms / (us_per_tile * 1e-6) / 1e12);
    }
    return 0;
}

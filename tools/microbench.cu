// microbench.cu -- B200 pipe measurements that decide K1 design choices (not product code).
//   M1  LDS.128 wavefront merging when lanes of different quarter-warps read identical addresses
//   M2  tcgen05.ld / tcgen05.st (32x32b) throughput: TMEM as per-lane storage
//   M3  SHFL throughput
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Run on the GPU box: ./tools/microbench   (prints cycles per warp-level instruction per SM)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- M1
template <int PATTERN>
__device__ __forceinline__ uint32_t lds_pattern(int lane) {
    switch (PATTERN) {
        case 0: return lane * 16;                                 // 32 distinct 16-byte words: 4 wavefronts
        case 1: return (lane & 15) * 16;                          // lanes l and l+16 identical
        case 2: return (lane & 7) * 16;                           // all four quarter-warps identical
        case 3: return (lane >> 1) * 16;                          // pairs inside a quarter-warp identical
        case 4: return ((lane & 3) + 4 * (lane >> 3)) * 16;       // quarter-warp = 2 x 4 positions
        case 5: return (lane & 15) * 16 + (lane >> 4) * 4096;     // half-warps in different buffers, same banks (conflict)
        default: return 0;
    }
}
template <int PATTERN>
__global__ void __launch_bounds__(1024, 1) lds128_kernel(unsigned long long *cycles, uint32_t *sink, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t base = smem_u32(sm) + lds_pattern<PATTERN>(lane) + ((threadIdx.x >> 5) & 3) * 2048;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const uint32_t ad = base + ((it & 7) << 11);
        uint32_t a0, b0, c0, d0, a1, b1, c1, d1, a2, b2, c2, d2, a3, b3, c3, d3;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(b0), "=r"(c0), "=r"(d0) : "r"(ad));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+512];" : "=r"(a1), "=r"(b1), "=r"(c1), "=r"(d1) : "r"(ad));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+1024];" : "=r"(a2), "=r"(b2), "=r"(c2), "=r"(d2) : "r"(ad));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+1536];" : "=r"(a3), "=r"(b3), "=r"(c3), "=r"(d3) : "r"(ad));
        s0 += a0 + b0 + c0 + d0; s1 += a1 + b1 + c1 + d1; s2 += a2 + b2 + c2 + d2; s3 += a3 + b3 + c3 + d3;
    }
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t acc = s0 ^ s1 ^ s2 ^ s3;
    if (acc == 0x12345678u) sink[0] = acc;
}

// ---------------------------------------------------------------- M3
__global__ void __launch_bounds__(256, 1) shfl_kernel(unsigned long long *cycles, uint32_t *sink, int iters) {
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = threadIdx.x * 31 + u;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __shfl_xor_sync(0xffffffffu, v[u], 1 + (u & 3));
    }
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t acc = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) acc ^= v[u];
    if (acc == 0x12345678u) sink[0] = acc;
}

// ---------------------------------------------------------------- M2
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
// MODE 0: loads only; 1: stores only; 2: load-modify-store round trip
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_kernel(unsigned long long *cycles, uint32_t *sink, int iters, int check) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = slot;
    const int warp = threadIdx.x >> 5;
    // lane quadrant = warp % 4; every warp sharing a quadrant gets its own 64-column window
    const uint32_t my = tbase + ((uint32_t)(warp & 3) << 21) + (uint32_t)(warp >> 2) * 64u;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = threadIdx.x * 64 + i;
    tmem_st32(my, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
            tmem_ld32(my, r);
            tmem_ld32(my + 32, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= r[it & 31];
        } else if (MODE == 1) {
            r[it & 31] ^= it;
            tmem_st32(my, r);
            tmem_st32(my + 32, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {
            tmem_ld32(my, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; i++) r[i] += 1;
            tmem_st32(my, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    const unsigned long long t1 = clock64();
    if (MODE == 2 && check) {
        tmem_ld32(my, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 32; i++) ok = ok && (r[i] == threadIdx.x * 64 + i + (uint32_t)iters);
        if (!ok) atomicAdd(&sink[1], 1u);
    }
#pragma unroll
    for (int i = 0; i < 32; i++) acc ^= r[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
    }
}

static double avg_cycles(unsigned long long *d, int n);

// ---------------------------------------------------------------- M4: STS.128 and mixed STS/LDS throughput
// MODE 0: 4 x STS.128; 1: 2 x STS.128 + 2 x LDS.128; 2: 4 x LDS.128 + 8 dependent-free DFMA per lane (pipe overlap)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) sts128_kernel(unsigned long long *cycles, uint32_t *sink, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t base = smem_u32(sm) + lane * 16 + ((threadIdx.x >> 5) & 3) * 2048;
    uint32_t s0 = threadIdx.x, s1 = 1, s2 = 2, s3 = 3;
    double f0 = 1.0, f1 = 1.1, f2 = 1.2, f3 = 1.3, f4 = 1.4, f5 = 1.5, f6 = 1.6, f7 = 1.7;
    const double m = 1.0000001, c = 1e-9;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const uint32_t ad = base + ((it & 7) << 11);
        if (MODE == 0 || MODE == 1) {
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ad), "r"(s0), "r"(s1), "r"(s2), "r"(s3));
            asm volatile("st.shared.v4.u32 [%0+512], {%1,%2,%3,%4};" ::"r"(ad), "r"(s1), "r"(s2), "r"(s3), "r"(s0));
        }
        if (MODE == 0) {
            asm volatile("st.shared.v4.u32 [%0+1024], {%1,%2,%3,%4};" ::"r"(ad), "r"(s2), "r"(s3), "r"(s0), "r"(s1));
            asm volatile("st.shared.v4.u32 [%0+1536], {%1,%2,%3,%4};" ::"r"(ad), "r"(s3), "r"(s0), "r"(s1), "r"(s2));
        }
        if (MODE == 1 || MODE == 2) {
            uint32_t a0, b0, c0, d0, a1, b1, c1, d1;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+1024];" : "=r"(a0), "=r"(b0), "=r"(c0), "=r"(d0) : "r"(ad));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+1536];" : "=r"(a1), "=r"(b1), "=r"(c1), "=r"(d1) : "r"(ad));
            s0 += a0 + b0; s1 += c0 + d0; s2 += a1 + b1; s3 += c1 + d1;
        }
        if (MODE == 2) {
            uint32_t a0, b0, c0, d0, a1, b1, c1, d1;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(b0), "=r"(c0), "=r"(d0) : "r"(ad));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+512];" : "=r"(a1), "=r"(b1), "=r"(c1), "=r"(d1) : "r"(ad));
            s0 += a0 + b0; s1 += c0 + d0; s2 += a1 + b1; s3 += c1 + d1;
            f0 = fma(f0, m, c); f1 = fma(f1, m, c); f2 = fma(f2, m, c); f3 = fma(f3, m, c);
            f4 = fma(f4, m, c); f5 = fma(f5, m, c); f6 = fma(f6, m, c); f7 = fma(f7, m, c);
        }
    }
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if ((s0 ^ s1 ^ s2 ^ s3) == 0x12345678u || f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 == 3.0) sink[0] = s0;
}
template <int MODE>
static int run_sts(unsigned long long *cyc, uint32_t *sink, const char *what) {
    const int iters = 4096;
    CK(cudaFuncSetAttribute(sts128_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int warps = 8; warps <= 32; warps *= 2) {
        sts128_kernel<MODE><<<148, warps * 32, 32768>>>(cyc, sink, iters);
        CK(cudaDeviceSynchronize());
        const double c = avg_cycles(cyc, 148);
        printf("M4 %s, %2d warps: %.2f cycles per warp-level 128-bit shared access per SM\n", what, warps, c / (iters * 4.0 * warps));
    }
    return 0;
}

static double avg_cycles(unsigned long long *d, int n) {
    unsigned long long h[1024];
    cudaMemcpy(h, d, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < n; i++) s += (double)h[i];
    return s / n;
}

template <int P>
static int run_lds(unsigned long long *cyc, uint32_t *sink, const char *what) {
    const int iters = 4096;
    CK(cudaFuncSetAttribute(lds128_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int warps = 4; warps <= 32; warps *= 2) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        lds128_kernel<P><<<148, warps * 32, 32768>>>(cyc, sink, iters);
        cudaEventRecord(e0);
        lds128_kernel<P><<<148, warps * 32, 32768>>>(cyc, sink, iters);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double c = avg_cycles(cyc, 148);
        printf("M1 LDS.128 pattern %d (%s), %d warps: %.2f cycles per warp-level LDS.128 per SM (clock64), kernel %.3f ms = %.0f cycles at 1.965 GHz (%d loads per warp)\n",
               P, what, warps, c / (iters * 4.0 * warps), ms, ms * 1.965e6, iters * 4);
    }
    return 0;
}

template <int MODE>
static int run_tmem(unsigned long long *cyc, uint32_t *sink, int warps, const char *what) {
    const int iters = 2048;
    tmem_kernel<MODE><<<148, warps * 32, 0>>>(cyc, sink, iters, 1);
    CK(cudaDeviceSynchronize());
    const double c = avg_cycles(cyc, 148);
    const double bytes = (MODE == 2 ? 2.0 : 2.0) * 32 * 32 * 4 * warps * iters;   // per SM
    printf("M2 TMEM %s, %2d warps: %.1f cycles/iter, %.1f bytes/cycle/SM\n", what, warps, c / iters, bytes / c);
    return 0;
}

int main() {
    unsigned long long *cyc;
    uint32_t *sink;
    CK(cudaMalloc(&cyc, 1024 * sizeof(unsigned long long)));
    CK(cudaMalloc(&sink, 16));
    CK(cudaMemset(sink, 0, 16));
    if (run_lds<0>(cyc, sink, "32 distinct addresses")) return 1;
    if (run_lds<1>(cyc, sink, "lanes l, l+16 identical")) return 1;
    if (run_lds<2>(cyc, sink, "all quarter-warps identical")) return 1;
    if (run_lds<3>(cyc, sink, "pairs inside a quarter-warp identical")) return 1;
    if (run_lds<4>(cyc, sink, "quarter-warp = 2 copies x 4 positions")) return 1;
    if (run_lds<5>(cyc, sink, "half-warps same banks, different rows")) return 1;
    {
        const int iters = 4096;
        shfl_kernel<<<148, 256>>>(cyc, sink, iters);
        CK(cudaDeviceSynchronize());
        printf("M3 SHFL: %.2f cycles per warp-level SHFL per SM (8 warps)\n", avg_cycles(cyc, 148) / (iters * 8.0 * 8.0));
    }
    if (run_sts<0>(cyc, sink, "4 x STS.128")) return 1;
    if (run_sts<1>(cyc, sink, "2 x STS.128 + 2 x LDS.128")) return 1;
    if (run_sts<2>(cyc, sink, "4 x LDS.128 + 8 DFMA")) return 1;
    for (int warps = 4; warps <= 16; warps *= 2) {
        if (run_tmem<0>(cyc, sink, warps, "ld 2 x (32x32b.x32) + wait")) return 1;
        if (run_tmem<1>(cyc, sink, warps, "st 2 x (32x32b.x32) + wait")) return 1;
        if (run_tmem<2>(cyc, sink, warps, "ld+wait, add, st+wait (x32)")) return 1;
    }
    uint32_t h[4];
    CK(cudaMemcpy(h, sink, 16, cudaMemcpyDeviceToHost));
    printf("TMEM round-trip mismatches: %u\n", h[1]);
    return 0;
}

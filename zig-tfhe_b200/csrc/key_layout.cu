// key_layout.cu -- one-time re-layout of CloudKey.bootstrapping_key (src/key.zig:61-65, 182-212) into
// the leaf order of the radix-8 transform, plus the FP64 roofline microbenchmark.
#include <cuda_runtime.h>

#include "exact_fft.cuh"
#include "kernels.cuh"
#include "negacyclic_fft.cuh"

namespace tfhe_b200 {

namespace {

// in : [chunks][2 (a,b)][1024] f64 = reference spectrum, re[0..512) | im[0..512), bin j (src/trlwe.zig:104-132)
// out: [chunks][ab][q0][t] cplx, leaf (q2,q1,q0) = reference bin (-(q2+8q1+64q0)) mod 512, scaled by 1/1024
//      (exact: collects the reference's x2 (fft.zig:356), x0.5 (trgsw.zig:174), x0.5 (fft.zig:380) and 1/512 (fft.zig:392))
__global__ void permute_bsk_kernel(const double *__restrict__ in, cplx *__restrict__ out, size_t polys) {
    const size_t poly = blockIdx.x;   // chunk * 2 + ab
    if (poly >= polys) return;
    const double *src = in + poly * kN;
    cplx *dst = out + poly * kHalfN;
    for (int s = threadIdx.x; s < kHalfN; s += blockDim.x) {
        const int q0 = s >> 6, t = s & 63;
        const int j = leaf_to_ref_bin(t >> 3, t & 7, q0);
        dst[s] = cplx{src[j] * (1.0 / 1024.0), src[kHalfN + j] * (1.0 / 1024.0)};
    }
}

// exact-mode device layout (exact_fft.cuh): out[chunk][ab][j0][t] = reference bin 64 j0 + 8 (t & 7) + (t >> 3), times 2^-10
__global__ void permute_bsk_exact_kernel(const double *__restrict__ in, cplx *__restrict__ out, size_t polys) {
    const size_t poly = blockIdx.x;   // chunk * 2 + ab
    if (poly >= polys) return;
    const double *src = in + poly * kN;
    cplx *dst = out + poly * kHalfN;
    for (int s = threadIdx.x; s < kHalfN; s += blockDim.x) {
        const int j = exact_bin(s >> 6, s & 63);
        dst[s] = cplx{src[j] * (1.0 / 1024.0), src[kHalfN + j] * (1.0 / 1024.0)};
    }
}

// lut.Generator.generateLookupTableFullAssign (src/lut/generator.zig:158-191) on the device, one CTA per table:
// raw[s] = table[x] for s in [divRound(x N, m), divRound((x+1) N, m)); rotated[i] = raw[(i + offset) mod N] with
// offset = divRound(N, 2m); the last `offset` coefficients negated; a = 0, b = rotated.
__global__ void build_testvec_kernel(const uint32_t *__restrict__ tables, int m, uint32_t *__restrict__ out) {
    const uint32_t *table = tables + (size_t)blockIdx.x * m;
    uint32_t *a = out + (size_t)blockIdx.x * 2 * kN, *b = a + kN;
    const int offset = (kN + m) / (2 * m);                        // divRound(N, 2m), generator.zig:253-255
    for (int i = threadIdx.x; i < kN; i += blockDim.x) {
        const int s = (i + offset) & (kN - 1);
        int x = (int)(((long long)s * m) / kN);                   // first guess, then settle on the divRound boundaries
        while (x + 1 < m && (((long long)(x + 1) * kN + m / 2) / m) <= s) x++;
        while (x > 0 && (((long long)x * kN + m / 2) / m) > s) x--;
        const uint32_t v = table[x];
        a[i] = 0u;
        b[i] = (i >= kN - offset) ? 0u - v : v;
    }
}

__global__ void fp64_peak_kernel(double *sink, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123456.789) sink[0] = r;   // never true; keeps the chain alive
}

}  // namespace

cudaError_t launch_permute_bsk(const double *ref_bsk, cplx *out, int n, int L, cudaStream_t s, uint64_t *launches) {
    const size_t polys = (size_t)n * 2 * L * 2;
    permute_bsk_kernel<<<(unsigned)polys, 256, 0, s>>>(ref_bsk, out, polys);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_permute_bsk_exact(const double *ref_bsk, cplx *out, int n, int L, cudaStream_t s, uint64_t *launches) {
    const size_t polys = (size_t)n * 2 * L * 2;
    permute_bsk_exact_kernel<<<(unsigned)polys, 256, 0, s>>>(ref_bsk, out, polys);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_build_testvec(const uint32_t *tables, int m, uint32_t *out, size_t count, cudaStream_t s, uint64_t *launches) {
    if (!count) return cudaSuccess;
    build_testvec_kernel<<<(unsigned)count, 256, 0, s>>>(tables, m, out);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t run_fp64_peak(int sm_count, cudaStream_t s, double *tflops, uint64_t *launches) {
    double *sink = nullptr;
    cudaError_t e = cudaMalloc(&sink, sizeof(double));
    if (e != cudaSuccess) return e;
    const int iters = 1 << 15, threads = 256, blocks = sm_count * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, threads, 0, s>>>(sink, 1 << 10, 0.999999, 1e-9);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, s);
        fp64_peak_kernel<<<blocks, threads, 0, s>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, s);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
        best = fmax(best, flops / (ms * 1e-3) / 1e12);
        if (launches) (*launches)++;
    }
    if (launches) (*launches)++;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return e;
}

}  // namespace tfhe_b200

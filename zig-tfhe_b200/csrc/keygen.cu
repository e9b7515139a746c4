// keygen.cu -- cloud-key generation on the device.
//
// Replaces (for a caller who holds the secret key on the host and wants the evaluation keys on the GPU):
//   key.genKeySwitchingKey             src/key.zig:148-172   -> keygen_ksk_kernel
//     tlwe.TLWELv0.encryptF64          src/tlwe.zig:34-50
//   key.genBootstrappingKey[WithRailgun]  src/key.zig:175-212 -> keygen_bsk_kernel
//     trgsw.TRGSWLv1.encryptTorus      src/trgsw.zig:35-71
//     trlwe.TRLWELv1.encryptF64        src/trlwe.zig:30-64   (a uniform, b = noise + a (*) s, negacyclic)
//     TRGSWLv1FFT.new / TRLWELv1FFT.new   src/trgsw.zig:80-91, src/trlwe.zig:108-132   (ifft1024 of a and b)
//   utils.NormalDist / gaussianF64 / f64ToTorus   src/utils.zig:28-33, 50-102
//
// The reference draws masks and noise from a clock-seeded xoshiro (src/utils.zig:16-22), i.e. its keys are not
// reproducible by construction; here every random word is a pure function of (seed, purpose, row, index) through
// Philox4x32-10, so any number of devices generate identical keys with no key traffic between them, and a key can
// be regenerated from its 8-byte seed.  Outputs land directly in the layouts K1 / K2 read (plus the reference
// layouts for export / the CPU oracle); nothing is staged through the host.
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.cuh"
#include "negacyclic_fft.cuh"

namespace tfhe_b200 {

namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// purposes (counter word 3): independent streams per key part
constexpr uint32_t kKskMask = 1, kKskNoise = 2, kBskMask = 3, kBskNoise = 4;

__device__ __forceinline__ uint4 rnd(uint64_t seed, uint32_t purpose, uint32_t row, uint32_t index) {
    return philox4x32_10(make_uint4(index, row, 0u, purpose), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// utils.f64ToTorus (src/utils.zig:28-33): @mod(d, 1) * 2^32, clamped to [0, 2^32 - 1], truncated
__device__ __forceinline__ uint32_t f64_to_torus(double d) {
    const double t = (d - floor(d)) * 4294967296.0;
    return (uint32_t)fmax(0.0, fmin(t, 4294967295.0));
}
// two N(0, alpha) samples by Box-Muller (utils.NormalDist.next, src/utils.zig:65-81) from one Philox block, as torus values
__device__ __forceinline__ void gaussian_pair(uint4 r, double alpha, uint32_t &t0, uint32_t &t1) {
    const double u1 = ((double)(r.x >> 5) * 67108864.0 + (double)(r.y >> 6) + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
    const double u2 = ((double)(r.z >> 5) * 67108864.0 + (double)(r.w >> 6)) * (1.0 / 9007199254740992.0);         // [0, 1)
    const double mag = alpha * sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    t0 = f64_to_torus(mag * c);
    t1 = f64_to_torus(mag * s);
}

struct KskArgs {
    const uint32_t *s0, *s1;
    uint64_t seed;
    double alpha;
    int n, basebit, iks_t, pitch;
    uint32_t *dev;      // [N][t][base-1][pitch]
    uint32_t *ref;      // [N*t*base][n+1] or nullptr
};

// one CTA per (i, j, k >= 1): TLWELv0.encryptF64(k * s1[i] / 2^((j+1) basebit), KSK_ALPHA, s0)
__global__ void __launch_bounds__(256) keygen_ksk_kernel(const KskArgs P) {
    __shared__ uint32_t part[8];
    const int base = 1 << P.basebit;
    const uint32_t row = blockIdx.x;                       // ((i * t + j) * (base - 1)) + (k - 1)
    const uint32_t k = row % (base - 1) + 1, ij = row / (base - 1);
    const uint32_t j = ij % P.iks_t, i = ij / P.iks_t;
    uint32_t *dev = P.dev + (size_t)row * P.pitch;
    uint32_t *ref = P.ref ? P.ref + ((size_t)base * P.iks_t * i + (size_t)base * j + k) * (P.n + 1) : nullptr;
    uint32_t inner = 0;
    for (int x4 = threadIdx.x; x4 * 4 < P.pitch; x4 += blockDim.x) {
        const uint4 r = rnd(P.seed, kKskMask, row, (uint32_t)x4);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int x = x4 * 4 + e;
            if (x < P.n) {
                inner += P.s0[x] * w[e];                   // tlwe.zig:39-43
                dev[x] = w[e];
                if (ref) ref[x] = w[e];
            } else if (x > P.n) {
                dev[x] = 0u;                               // padding columns of the device layout
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) inner += __shfl_xor_sync(0xffffffffu, inner, s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = inner;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) inner += part[w];
        uint32_t e0, e1;
        gaussian_pair(rnd(P.seed, kKskNoise, row, 0u), P.alpha, e0, e1);
        const double p = ((double)k * (double)P.s1[i]) / (double)(1u << ((j + 1) * P.basebit));   // key.zig:164
        const uint32_t b = inner + (e0 + f64_to_torus(p));                                       // utils.zig:85-102, tlwe.zig:46-47
        dev[P.n] = b;
        if (ref) ref[P.n] = b;
    }
}

struct BskArgs {
    const uint32_t *s0, *s1;
    uint64_t seed;
    double alpha;
    int n, L, bgbit;
    const cplx *tw2, *tw3;
    cplx *dev;          // [n*2L] chunks of [ab][q0][t]
    double *ref;        // [n][2L][2][N] reference spectra (kept on the device for exact mode, exported on request)
};

// forward transform of one polynomial (signed coefficients in shared memory) by a 64-thread group; leaf values out in v
__device__ __forceinline__ void group_forward(const uint32_t *poly, cplx *x1, cplx *x2, const cplx *tw2, const cplx *tw3, int t, int barid,
                                              cplx (&v)[8]) {
    const int hi = t >> 3, lo = t & 7;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int e = 64 * p + 8 * lo + hi;
        v[p] = cplx{(double)(int32_t)poly[e], (double)(int32_t)poly[e + kHalfN]};     // ifft1024 reads i32 -> f64 (fft.zig:316-327)
    }
    fwd_pass1(v);
#pragma unroll
    for (int q = 0; q < 8; q++) x1[x1_slot(hi, q, lo)] = v[q];
    asm volatile("bar.sync %0, 64;" ::"r"(barid) : "memory");
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = x1[x1_slot(hi, lo, q)];
    fwd_pass(v, tw2 + lo, 8);
#pragma unroll
    for (int q = 0; q < 8; q++) x2[x2_slot(lo, q, hi)] = v[q];
    asm volatile("bar.sync %0, 64;" ::"r"(barid) : "memory");
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = x2[x2_slot(hi, lo, q)];
    fwd_pass(v, tw3 + t, 64);
}

// one CTA (128 threads) per TRLWE row (i, r) of the bootstrapping key
__global__ void __launch_bounds__(128) keygen_bsk_kernel(const BskArgs P) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t *a = reinterpret_cast<uint32_t *>(smem);           // [N]
    uint32_t *b = a + kN;                                        // [N]
    uint16_t *ones = reinterpret_cast<uint16_t *>(b + kN);       // indices i with s1[i] = 1
    int *n_ones = reinterpret_cast<int *>(ones + kN);
    cplx *xbuf = reinterpret_cast<cplx *>(smem + 2 * kN * 4 + kN * 2 + 16);   // two groups x (X1 + X2)
    const int tid = threadIdx.x;
    const uint32_t row = blockIdx.x;                             // i * 2L + r
    const uint32_t i = row / (2 * P.L), r = row % (2 * P.L);
    if (tid == 0) *n_ones = 0;
    // a uniform, b = gaussian noise with mean 0 (trlwe.zig:39-52)
    for (int x4 = tid; x4 < kN / 4; x4 += blockDim.x) {
        const uint4 u = rnd(P.seed, kBskMask, row, (uint32_t)x4);
        a[4 * x4] = u.x; a[4 * x4 + 1] = u.y; a[4 * x4 + 2] = u.z; a[4 * x4 + 3] = u.w;
    }
    for (int x2 = tid; x2 < kN / 2; x2 += blockDim.x) gaussian_pair(rnd(P.seed, kBskNoise, row, (uint32_t)x2), P.alpha, b[2 * x2], b[2 * x2 + 1]);
    __syncthreads();
    for (int x = tid; x < kN; x += blockDim.x)
        if (P.s1[x] & 1u) ones[atomicAdd(n_ones, 1)] = (uint16_t)x;
    __syncthreads();
    // b += a (*) s1 in Z[X]/(X^N + 1) (trlwe.zig:54-61); s1 is binary (key.zig:23-58), so the product is a signed sum
    const int cnt = *n_ones;
    uint32_t acc[kN / 128];
#pragma unroll
    for (int m = 0; m < kN / 128; m++) acc[m] = 0u;
    for (int q = 0; q < cnt; q++) {
        const int s = ones[q];
#pragma unroll
        for (int m = 0; m < kN / 128; m++) {
            const int j = tid + 128 * m;
            const uint32_t v = a[(j - s) & (kN - 1)];
            acc[m] += (j >= s) ? v : 0u - v;
        }
    }
#pragma unroll
    for (int m = 0; m < kN / 128; m++) b[tid + 128 * m] += acc[m];
    __syncthreads();
    // gadget term on the constant coefficient (trgsw.zig:43-51, 64-68): s0[i] * BG^-(l+1)
    if (tid == 0) {
        const uint32_t l = r % P.L;
        const uint32_t val = P.s0[i] * f64_to_torus(exp2(-(double)((l + 1) * P.bgbit)));
        if (r < (uint32_t)P.L) a[0] += val;
        else b[0] += val;
    }
    __syncthreads();
    // spectra: group 0 transforms a, group 1 transforms b
    const int g = tid >> 6, t = tid & 63;
    cplx v[8];
    cplx *x1 = xbuf + (size_t)g * (kX1Slots + kX2Slots);
    group_forward(g ? b : a, x1, x1 + kX1Slots, P.tw2, P.tw3, t, 1 + g, v);
    cplx *dev = P.dev + (size_t)row * kBskChunkCplx;
    double *ref = P.ref + ((size_t)row * 2 + g) * kN;
#pragma unroll
    for (int q0 = 0; q0 < 8; q0++) {
        // leaf value Z: the reference's ifft1024 bin is 2 Z (fft.zig:356); the device layout holds bin / 1024 (key_layout.cu)
        dev[bsk_slot(g, q0, t)] = cplx{v[q0].re * (1.0 / 512.0), v[q0].im * (1.0 / 512.0)};
        const int j = leaf_to_ref_bin(t >> 3, t & 7, q0);
        ref[j] = 2.0 * v[q0].re;
        ref[kHalfN + j] = 2.0 * v[q0].im;
    }
}

}  // namespace

cudaError_t launch_keygen_ksk(const uint32_t *s0, const uint32_t *s1, uint64_t seed, double alpha, int n, int basebit, int iks_t, int pitch,
                              uint32_t *dev, uint32_t *ref, cudaStream_t s, uint64_t *launches) {
    KskArgs A{s0, s1, seed, alpha, n, basebit, iks_t, pitch, dev, ref};
    const unsigned rows = (unsigned)kN * iks_t * ((1u << basebit) - 1u);
    keygen_ksk_kernel<<<rows, 256, 0, s>>>(A);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_keygen_bsk(const uint32_t *s0, const uint32_t *s1, uint64_t seed, double alpha, int n, int L, int bgbit, const cplx *tw2,
                              const cplx *tw3, cplx *dev, double *ref, cudaStream_t s, uint64_t *launches) {
    BskArgs A{s0, s1, seed, alpha, n, L, bgbit, tw2, tw3, dev, ref};
    const size_t smem = 2 * kN * 4 + kN * 2 + 16 + (size_t)2 * (kX1Slots + kX2Slots) * sizeof(cplx);
    cudaError_t e = cudaFuncSetAttribute(keygen_bsk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    keygen_bsk_kernel<<<(unsigned)(n * 2 * L), 128, smem, s>>>(A);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace tfhe_b200

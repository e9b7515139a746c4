// blind_rotate_exact.cu -- K1x: blind rotation in EXACT mode.
//
// Replays the reference's floating-point DAG operation for operation, so that the accumulator is
// bit-identical to zig-tfhe's even on the large-digit parameter sets (UINT1-8, BGBIT 10..23) where
// the FP64 external product is no longer an exact integer computation and any other transform
// (such as the radix-8 FMA one of blind_rotate.cu) rounds differently:
//   ifft1024  : twist (mul, mul, sub / mul, mul, add, no FMA)            src/fft.zig:293-334
//               bit reversal + radix-2 DIT, twiddles = the values the serial recurrence
//               w <- w * w_len produces (tabulated by the host with the same recurrence)   src/fft.zig:582-669
//               x2                                                         src/fft.zig:339-357
//   fmaInFd1024: res += (a_re*b_re - a_im*b_im) * 0.5, rows in order       src/trgsw.zig:139-142, 157-189
//   fft1024   : x0.5, inverse radix-2, untwist, x(1/512), @round (half away from zero),
//               i64 -> truncating i32                                      src/fft.zig:370-443
// The bootstrapping key is used in the reference's own layout and scaling (no permutation).
//
// One CTA of 256 threads per ciphertext, one radix-2 butterfly per thread per stage.  This is the
// parity path, not the throughput path: it exists so that "GPU == reference" can be demonstrated on
// every parameter set; the fast kernel is the product path wherever it is bit-identical.
#include <cuda_runtime.h>

#include "br_common.cuh"
#include "kernels.cuh"

namespace tfhe_b200 {

namespace {

constexpr int kThreads = 256;
constexpr int kTabStride = 512;   // exact_tables: twist_re, twist_im, fwd_re, fwd_im, inv_re, inv_im (512 doubles each)

// radix-2 DIT over 512 complex points held as split re/im in shared memory (fft.zig:589-618);
// input already bit-reversed.  tw_re/tw_im: stage s occupies [2^s - 1, 2^(s+1) - 1).
__device__ __forceinline__ void radix2_stages(double *re, double *im, const double *tw_re, const double *tw_im, int j) {
#pragma unroll 1
    for (int s = 0; s < 9; s++) {
        const int half = 1 << s;
        const int pos = j & (half - 1);
        const int i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
        const double w_re = tw_re[half - 1 + pos], w_im = tw_im[half - 1 + pos];
        const double u_re = re[i0], u_im = im[i0], d_re = re[i1], d_im = im[i1];
        // Complex.mul(self = data, other = w), fft.zig:51-55
        const double v_re = __dadd_rn(__dmul_rn(d_re, w_re), -__dmul_rn(d_im, w_im));
        const double v_im = __dadd_rn(__dmul_rn(d_re, w_im), __dmul_rn(d_im, w_re));
        re[i0] = __dadd_rn(u_re, v_re); im[i0] = __dadd_rn(u_im, v_im);       // fft.zig:605
        re[i1] = __dadd_rn(u_re, -v_re); im[i1] = __dadd_rn(u_im, -v_im);     // fft.zig:606
        __syncthreads();
    }
}

__device__ __forceinline__ int brev9(int i) { return (int)(__brev((unsigned)i) >> 23); }

template <bool MARGIN>
__global__ void __launch_bounds__(kThreads, 2) blind_rotate_exact_kernel(const BrArgs P, const double *__restrict__ tables,
                                                                        const double *__restrict__ bsk_ref) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tab = reinterpret_cast<double *>(smem_raw);             // 6 * 512 doubles
    double *re = tab + 6 * kTabStride, *im = re + kHalfN;          // work buffer
    double *oa = im + kHalfN, *ob = oa + kN;                       // output spectra (re | im split like the reference)
    uint32_t *acc_a = reinterpret_cast<uint32_t *>(ob + kN), *acc_b = acc_a + kN;
    uint32_t *diff = acc_b + kN;                                   // (rot - acc + offset) of the current polynomial
    uint16_t *atil = reinterpret_cast<uint16_t *>(diff + kN);
    const int n = P.n, L = P.L, bgbit = P.bgbit;
    const int j = threadIdx.x;
    const size_t ct = (size_t)P.ct_base + blockIdx.x;

    for (int i = j; i < 6 * kTabStride; i += kThreads) tab[i] = tables[i];
    const double *twist_re = tab, *twist_im = tab + kTabStride;
    const double *fwd_re = tab + 2 * kTabStride, *fwd_im = tab + 3 * kTabStride;
    const double *inv_re = tab + 4 * kTabStride, *inv_im = tab + 5 * kTabStride;
    {
        const GateOperands go = gate_operands(P, ct, n);
        const int op = go.op;
        for (int i = j; i <= n; i += kThreads) {
            uint32_t lin = gate_linear_signed(go, i);
            if (i == n) lin += gate_constant(op);
            const uint32_t m = (uint32_t)(((unsigned long long)lin + (1u << 20)) >> 21);
            atil[i] = (uint16_t)((i == n) ? (2 * kN - m) : m);
        }
    }
    __syncthreads();
    {
        const int btil = atil[n];
        const uint32_t *tv = P.testvec ? P.testvec + (P.tv_per_item ? ct * (size_t)(2 * kN) : 0) : nullptr;
        for (int e = j; e < kN; e += kThreads) {
            const int u = (e - btil) & (2 * kN - 1);
            const uint32_t va = tv ? tv[u & (kN - 1)] : 0u;
            const uint32_t vb = tv ? tv[kN + (u & (kN - 1))] : 0x20000000u;
            acc_a[e] = (u & kN) ? 0u - va : va;
            acc_b[e] = (u & kN) ? 0u - vb : vb;
        }
    }
    __syncthreads();

    const uint32_t mask = (1u << bgbit) - 1u, half_bg = 1u << (bgbit - 1);
    double margin = 0.0;
    for (int i = 0; i < n; i++) {
        const int at = atil[i];
        for (int e = j; e < kN; e += kThreads) { oa[e] = 0.0; ob[e] = 0.0; }
        for (int h = 0; h < 2; h++) {
            const uint32_t *accp = h ? acc_b : acc_a;
            for (int e = j; e < kN; e += kThreads) {                    // cmux difference, trgsw.zig:270-273 + 208-209
                const int u = (e - at) & (2 * kN - 1);
                const uint32_t v = accp[u & (kN - 1)];
                diff[e] = ((u & kN) ? 0u - v : v) - accp[e] + P.offset;
            }
            __syncthreads();
            for (int l = 0; l < L; l++) {
                const int sh = 32 - (l + 1) * bgbit;
                for (int k = j; k < kHalfN; k += kThreads) {            // fold + twist, fft.zig:297-334
                    const double in_re = (double)(int32_t)(((diff[k] >> sh) & mask) - half_bg);
                    const double in_im = (double)(int32_t)(((diff[k + kHalfN] >> sh) & mask) - half_bg);
                    const double w_re = twist_re[k], w_im = twist_im[k];
                    const int dst = brev9(k);                           // bitReverseRadix2, fft.zig:647-669
                    re[dst] = __dadd_rn(__dmul_rn(in_re, w_re), -__dmul_rn(in_im, w_im));
                    im[dst] = __dadd_rn(__dmul_rn(in_re, w_im), __dmul_rn(in_im, w_re));
                }
                __syncthreads();
                radix2_stages(re, im, fwd_re, fwd_im, j);
                const double *ba = bsk_ref + (((size_t)i * 2 * L + h * L + l) * 2 + 0) * kN;
                const double *bb = ba + kN;
                for (int k = j; k < kHalfN; k += kThreads) {            // fmaInFd1024, trgsw.zig:174-181
                    const double a_re = __dmul_rn(re[k], 2.0), a_im = __dmul_rn(im[k], 2.0);   // fft.zig:356-357
                    {
                        const double b_re = ba[k], b_im = ba[k + kHalfN];
                        oa[k] = __dadd_rn(oa[k], __dmul_rn(__dadd_rn(__dmul_rn(a_re, b_re), -__dmul_rn(a_im, b_im)), 0.5));
                        oa[k + kHalfN] = __dadd_rn(oa[k + kHalfN], __dmul_rn(__dadd_rn(__dmul_rn(a_re, b_im), __dmul_rn(a_im, b_re)), 0.5));
                    }
                    {
                        const double b_re = bb[k], b_im = bb[k + kHalfN];
                        ob[k] = __dadd_rn(ob[k], __dmul_rn(__dadd_rn(__dmul_rn(a_re, b_re), -__dmul_rn(a_im, b_im)), 0.5));
                        ob[k + kHalfN] = __dadd_rn(ob[k + kHalfN], __dmul_rn(__dadd_rn(__dmul_rn(a_re, b_im), __dmul_rn(a_im, b_re)), 0.5));
                    }
                }
                __syncthreads();
            }
        }
        for (int h = 0; h < 2; h++) {                                    // fft1024, fft.zig:370-443
            const double *o = h ? ob : oa;
            uint32_t *accp = h ? acc_b : acc_a;
            for (int k = j; k < kHalfN; k += kThreads) {
                const int dst = brev9(k);
                re[dst] = __dmul_rn(o[k], 0.5);
                im[dst] = __dmul_rn(o[k + kHalfN], 0.5);
            }
            __syncthreads();
            radix2_stages(re, im, inv_re, inv_im, j);
            for (int k = j; k < kHalfN; k += kThreads) {
                const double w_re = twist_re[k], w_im = twist_im[k];
                const double f_re = re[k], f_im = im[k];
                const double t_re = __dmul_rn(__dadd_rn(__dmul_rn(f_re, w_re), __dmul_rn(f_im, w_im)), 1.0 / 512.0);   // fft.zig:416
                const double t_im = __dmul_rn(__dadd_rn(__dmul_rn(f_im, w_re), -__dmul_rn(f_re, w_im)), 1.0 / 512.0);  // fft.zig:417
                const double r_re = round(t_re), r_im = round(t_im);    // @round: half away from zero
                if (MARGIN) margin = fmax(margin, fmax(fabs(t_re - r_re), fabs(t_im - r_im)));
                accp[k] += (uint32_t)(unsigned long long)__double2ll_rz(r_re);       // cmux add-back, trgsw.zig:278-281
                accp[k + kHalfN] += (uint32_t)(unsigned long long)__double2ll_rz(r_im);
            }
            __syncthreads();
        }
    }

    if (P.out_trlwe) {
        uint32_t *o = P.out_trlwe + ct * (size_t)(2 * kN);
        for (int e = j; e < kN; e += kThreads) { o[e] = acc_a[e]; o[kN + e] = acc_b[e]; }
    }
    if (P.out_lv1) {
        uint32_t *o = P.out_lv1 + ct * (size_t)(kN + 1);
        for (int e = j; e <= kN; e += kThreads) o[e] = (e == 0) ? acc_a[0] : (e == kN) ? acc_b[0] : 0u - acc_a[kN - e];
    }
    if (MARGIN && P.margin_bits) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) margin = fmax(margin, __shfl_xor_sync(0xffffffffu, margin, s));
        if ((j & 31) == 0) atomicMax(P.margin_bits, (unsigned long long)__double_as_longlong(margin));
    }
}

}  // namespace

// exact_tables: device copy of make_exact_tables() (host_tables.h); bsk_ref: CloudKey.bootstrapping_key as loaded
cudaError_t launch_blind_rotate_exact(const BrArgs &a, const double *exact_tables, const double *bsk_ref, bool track_margin,
                                      cudaStream_t s, uint64_t *launches) {
    if (a.B == 0) return cudaSuccess;
    const size_t smem = (6 * kTabStride + 2 * kHalfN + 2 * kN) * sizeof(double) + 3 * kN * sizeof(uint32_t) + (((size_t)a.n + 1) * 2 + 15) / 16 * 16;
    auto k0 = blind_rotate_exact_kernel<false>;
    auto k1 = blind_rotate_exact_kernel<true>;
    cudaError_t e = cudaFuncSetAttribute(track_margin ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (launches) (*launches)++;
    if (track_margin) k1<<<a.B, kThreads, smem, s>>>(a, exact_tables, bsk_ref);
    else k0<<<a.B, kThreads, smem, s>>>(a, exact_tables, bsk_ref);
    return cudaGetLastError();
}

}  // namespace tfhe_b200

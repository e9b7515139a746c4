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
//
// Two kernels produce those bits:
//  * blind_rotate_exact_rb_kernel ("register-blocked", the default): the same radix-2 butterflies, three stages at a
//    time on 8 points in one thread's registers (exact_fft.cuh), 64 threads per ciphertext, KCT ciphertexts per CTA
//    sharing the key ring -- the geometry of the fast kernel (same exchanges, accumulator layout, cp.async.bulk ring),
//    only the arithmetic inside each pass differs.  The key is read in the exact chunk layout (permuted, times 2^-10).
//  * blind_rotate_exact_kernel (round 1, "legacy"): one CTA of 256 threads per ciphertext, one butterfly per thread per
//    stage through shared memory, key in the reference's own layout.  Kept for A/B runs (tuning key "exact_legacy") and
//    as the fallback if the host libm's stage tables are not conjugate-symmetric.
#include <cuda_runtime.h>

#include "br_common.cuh"
#include "br_ring.cuh"
#include "exact_fft.cuh"
#include "host_tables.h"
#include "kernels.cuh"

#include <vector>

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
            const uint32_t m = mod_switch_2n(lin, P.ms_shift);
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


// ---------------------------------------------------------------------------------------------------
// Register-blocked exact kernel.
struct ExactConsts {
    cplx twa2, twa4, twa5, twa6;     // the pass-A twiddles that are not (1, 0): stage-table entries 2, 4, 5, 6 (kernel
                                     // parameters live in the constant bank, so they cost no registers)
};

// Per-thread stage twiddles of passes B and C: resident in registers (2 x 28), or -- DENSE, six ciphertexts per CTA at 168
// registers -- parked in 64 tensor-memory columns of the thread's own lane and fetched by one tcgen05.ld before each pass
// (same scheme as the fast kernel's "twt", blind_rotate.cu).
template <bool DENSE>
struct ExTwSrc {
    ExTw b, c;            // !DENSE
    uint32_t taddr;       // DENSE: pass B in columns [0, 32), pass C in [32, 64)
    __device__ __forceinline__ ExTw get(int which) const {
        if (!DENSE) return which ? c : b;
        uint32_t r[32];
        tmem_ld32(taddr + 32u * (uint32_t)which, r);
        tmem_wait_ld();
        ExTw w;
        auto cp = [&](int k) { return cplx{__hiloint2double((int)r[4 * k + 1], (int)r[4 * k]), __hiloint2double((int)r[4 * k + 3], (int)r[4 * k + 2])}; };
        w.wa = cp(0); w.wb[0] = cp(1); w.wb[1] = cp(2); w.wc[0] = cp(3); w.wc[1] = cp(4); w.wc[2] = cp(5); w.wc[3] = cp(6);
        return w;
    }
};
__device__ __forceinline__ void ex_tw_park(uint32_t taddr, const ExTw &w) {
    uint32_t r[32];
    const cplx v[7] = {w.wa, w.wb[0], w.wb[1], w.wc[0], w.wc[1], w.wc[2], w.wc[3]};
#pragma unroll
    for (int k = 0; k < 7; k++) {
        r[4 * k] = (uint32_t)__double2loint(v[k].re); r[4 * k + 1] = (uint32_t)__double2hiint(v[k].re);
        r[4 * k + 2] = (uint32_t)__double2loint(v[k].im); r[4 * k + 3] = (uint32_t)__double2hiint(v[k].im);
    }
#pragma unroll
    for (int k = 28; k < 32; k++) r[k] = 0u;
    tmem_st32(taddr, r);
}

// DENSE also lays X1 over this warp's rows of the X2 buffer the transform does not use for its own X2 exchange (x1a_slot):
// every exact transform runs A -> X1 -> B -> X2 -> C, its X2 reads are partitioned by warp, and the other warp writes that
// buffer again only one transform later, after the group barrier this warp reaches with X1 long finished.
template <bool CONJ, bool DENSE>
__device__ __forceinline__ void ex_transform(cplx (&v)[8], Xbuf &xb, const ExTwSrc<DENSE> &tw, const ExactConsts &kc, int hi, int lo, int barid,
                                             Producer &pr) {
    {
        const cplx twa[kExactPassATw] = {cplx{1.0, 0.0}, cplx{1.0, 0.0}, kc.twa2, cplx{1.0, 0.0}, kc.twa4, kc.twa5, kc.twa6};
        ex_pass_a<CONJ>(v, twa);
    }
    cplx *x1 = DENSE ? xb.x2 + (xb.flip ^ kX2Slots) : xb.x1;
#pragma unroll
    for (int q = 0; q < 8; q++) x1[DENSE ? x1a_slot(hi, q, lo) : x1_slot(hi, q, lo)] = v[q];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = x1[DENSE ? x1a_slot(hi, lo, q) : x1_slot(hi, lo, q)];
    {
        const ExTw w = tw.get(0);
        ex_pass<CONJ>(v, w.wa, w.wb, w.wc);
    }
    if (!DENSE) producer_poll(pr);
    cplx *x2 = xb.x2 + xb.flip;      // double-buffered: one named barrier per transform (see Xbuf)
    xb.flip ^= kX2Slots;
#pragma unroll
    for (int q = 0; q < 8; q++) x2[x2_slot(lo, q, hi)] = v[q];
    bar_sync(barid, kGroupThreads);
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = x2[x2_slot(hi, lo, q)];
    {
        const ExTw w = tw.get(1);
        ex_pass<CONJ>(v, w.wa, w.wb, w.wc);
    }
}

constexpr int kExStages = 3;
__host__ __device__ constexpr int ex_group_bytes(int n, bool dense) { return 2 * kN * 4 + (dense ? 0 : kX1Slots * 16) + 2 * kX2Slots * 16 + align16((n + 1) * 2); }
__host__ __device__ constexpr int ex_fixed_bytes() { return kExStages * kBskChunkBytes + 112 + kExactSharedTabCplx * 16; }

// LT / BGT > 0: gadget length and digit width as compile-time constants (digit loop unrolled, shifts and masks folded), as in the
// fast kernel, where that instantiation is worth 6 % (profiles/r02_k1_ring.log); instantiated for L = 1 / BGBIT = 22 (UINT4 ... UINT8):
// 126.7 k -> 136.9 k UINT4 bootstraps/s (profiles/r02_k1x_dense.log).
template <int KCT, bool MARGIN, int LT = 0, int BGT = 0>
__global__ void __launch_bounds__(KCT * kGroupThreads, 1)
    blind_rotate_exact_rb_kernel(const BrArgs P, const double *__restrict__ tables, const cplx *__restrict__ bsk_x,
                                 const cplx *__restrict__ shared_tab, const ExactConsts kc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ptr = smem_raw;
    cplx *bsk_ring = reinterpret_cast<cplx *>(ptr);
    ptr += kExStages * kBskChunkBytes;
    constexpr bool DENSE = KCT > 4;      // six ciphertexts per CTA: tensor-memory twiddles, X1 over X2, last-arriver key ring
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(ptr);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ptr + 64);
    uint32_t *ring_cnt = reinterpret_cast<uint32_t *>(ptr + 80);   // [kMaxStages] releases per stage (last-arriver refill)
    ptr += 112;
    cplx *twist = reinterpret_cast<cplx *>(ptr);     // [512] in acc_pos order
    ptr += kExactSharedTabCplx * 16;
    const int n = P.n, L = LT > 0 ? LT : P.L, bgbit = BGT > 0 ? BGT : P.bgbit;
    const int tid = threadIdx.x;
    const int first_ct = blockIdx.x * KCT;
    const int n_active = min(KCT, (int)P.B - first_ct);
    if (tid == 0) {
        for (int s = 0; s < kExStages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], n_active * 2);   // one arrival per consumer warp
            ring_cnt[s] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (DENSE && tid < 32) {
        tmem_alloc(tmem_slot, 512);     // 96 columns per warp (stage twiddles + twist), three warps per lane quadrant
        tmem_fence_before_sync();
    }
    for (int j = tid; j < kExactSharedTabCplx; j += KCT * kGroupThreads) twist[j] = shared_tab[j];
    __syncthreads();
    uint32_t tmem_base = 0;
    if (DENSE) {
        tmem_fence_after_sync();
        tmem_base = *tmem_slot;
    }
    const int g = tid >> 6, t = tid & 63;
    if (g >= n_active) return;     // whole warps (warp 0 is always active: it frees the tensor-memory allocation at the end)
    const int hi = t >> 3, lo = t & 7;
    const int barid = 1 + g;
    unsigned char *gb = ptr + (size_t)g * ex_group_bytes(n, DENSE);
    uint32_t *acc_a = reinterpret_cast<uint32_t *>(gb), *acc_b = acc_a + kN;
    constexpr int kX1B = DENSE ? 0 : kX1Slots * 16;
    Xbuf xb;
    xb.x1 = reinterpret_cast<cplx *>(gb + 2 * kN * 4);
    xb.x2 = reinterpret_cast<cplx *>(gb + 2 * kN * 4 + kX1B);
    xb.flip = 0;
    uint16_t *atil = reinterpret_cast<uint16_t *>(gb + 2 * kN * 4 + kX1B + 2 * kX2Slots * 16);
    const size_t ct = (size_t)P.ct_base + first_ct + g;

    // per-thread stage twiddles (forward table; the inverse direction uses their exact conjugates)
    ExTwSrc<DENSE> tw;
    if (DENSE) {
        tw.taddr = tmem_base + ((uint32_t)((tid >> 5) & 3) << 21) + (uint32_t)(tid >> 7) * 96u;
        {   // this thread's eight twist values (coefficients 64 p + 8 lo + hi) in columns [64, 96) of its window
            uint32_t r[32];
#pragma unroll
            for (int p = 0; p < 8; p++) {
                const cplx w = twist[64 * p + t];
                r[4 * p] = (uint32_t)__double2loint(w.re); r[4 * p + 1] = (uint32_t)__double2hiint(w.re);
                r[4 * p + 2] = (uint32_t)__double2loint(w.im); r[4 * p + 3] = (uint32_t)__double2hiint(w.im);
            }
            tmem_st32(tw.taddr + 64u, r);
        }
        ex_tw_park(tw.taddr, ex_twiddles_b(tables + 2 * kExactTabStride, tables + 3 * kExactTabStride, lo));
        ex_tw_park(tw.taddr + 32u, ex_twiddles_c(tables + 2 * kExactTabStride, tables + 3 * kExactTabStride, 8 * lo + hi));
        tmem_wait_st();
    } else {
        tw.b = ex_twiddles_b(tables + 2 * kExactTabStride, tables + 3 * kExactTabStride, lo);
        tw.c = ex_twiddles_c(tables + 2 * kExactTabStride, tables + 3 * kExactTabStride, 8 * lo + hi);
    }

    {   // gate linear part (gates.zig:48-121) + modulus switch (trgsw.zig:297,312)
        const GateOperands go = gate_operands(P, ct, n);
        const int op = go.op;
        for (int i = t; i <= n; i += kGroupThreads) {
            uint32_t lin = gate_linear_signed(go, i);
            if (i == n) lin += gate_constant(op);
            const uint32_t m = mod_switch_2n(lin, P.ms_shift);
            atil[i] = (uint16_t)((i == n) ? (2 * kN - m) : m);
        }
    }
    bar_sync(barid, kGroupThreads);
    {   // acc = X^btil * testvec (trgsw.zig:300-306), acc_pos order
        const int btil = atil[n];
        const uint32_t *tv = P.testvec ? P.testvec + (P.tv_per_item ? ct * (size_t)(2 * kN) : 0) : nullptr;
        for (int j = t; j < kN; j += kGroupThreads) {
            const int u = (j - btil) & (2 * kN - 1);
            const uint32_t va = tv ? tv[u & (kN - 1)] : 0u;
            const uint32_t vb = tv ? tv[kN + (u & (kN - 1))] : 0x20000000u;
            acc_a[acc_pos(j)] = (u & kN) ? 0u - va : va;
            acc_b[acc_pos(j)] = (u & kN) ? 0u - vb : vb;
        }
    }
    bar_sync(barid, kGroupThreads);

    const uint32_t mask = (1u << bgbit) - 1u, half_bg = 1u << (bgbit - 1);
    const uint32_t offset = P.offset;
    int stage = 0;
    uint32_t phase = 0;
    double margin = 0.0;
    Producer pr;
    pr.ring = bsk_ring; pr.full_bar = full_bar; pr.empty_bar = empty_bar; pr.stages = kExStages;
    pr.ring_cnt = ring_cnt; pr.bsk = bsk_x; pr.total = (uint32_t)(n * 2 * L); pr.ring_warps = (uint32_t)(n_active * 2);
    if (DENSE) {     // last-arriver refill (br_ring.cuh): thread 0 fills the ring once, the consumers issue every later bulk copy
        pr.active = false; pr.remaining = 0;
        if (tid == 0) {
            const uint64_t policy = l2_policy_evict_last();
            for (int s = 0; s < kExStages && s < n * 2 * L; s++) {
                mbar_arrive_expect_tx(&full_bar[s], kBskChunkBytes);
                bulk_g2s(bsk_ring + s * kBskChunkCplx, bsk_x + (size_t)s * kBskChunkCplx, kBskChunkBytes, &full_bar[s], policy);
            }
        }
    } else {
        pr.src = bsk_x;
        pr.remaining = n * 2 * L; pr.issued = 0; pr.stage = 0; pr.phase = 0;
        pr.active = tid == 0;
        pr.policy = pr.active ? l2_policy_evict_last() : 0;
#pragma unroll
        for (int s = 0; s < kExStages; s++) producer_poll(pr);
    }
    __nanosleep((blockIdx.x % 41u) * 128u);   // de-synchronise the CTAs of a wave (see blind_rotate.cu)

    for (int i = 0; i < n; i++) {
        const int at = atil[i];
        cplx oa[8], ob[8];
#pragma unroll
        for (int q = 0; q < 8; q++) { oa[q] = cplx{0.0, 0.0}; ob[q] = cplx{0.0, 0.0}; }
// the two polynomials unrolled where there is one digit each (L = 1 instantiation: 144.2 k -> 149.3 k UINT4 blind rotations/s; the same
// unrolling costs the fast kernel 5.6 %, profiles/r02_k1_ring.log)
#pragma unroll (LT == 1 ? 2 : 1)
        for (int h = 0; h < 2; h++) {
            const uint32_t *accp = h ? acc_b : acc_a;
            uint32_t d[16];
            load_rot_diffs(d, accp, at, offset, hi, lo);
#pragma unroll (LT > 0 ? LT : 1)
            for (int l = 0; l < L; l++) {
                const int sh = 32 - (l + 1) * bgbit;
                cplx v[8];
                uint32_t twr[32];
                if (DENSE) {      // twist values from tensor memory (shared-memory table otherwise)
                    tmem_ld32(tw.taddr + 64u, twr);
                    tmem_wait_ld();
                }
#pragma unroll
                for (int p = 0; p < 8; p++) {     // decomposition digit (trgsw.zig:208-217), fold + twist (fft.zig:297-334)
                    const double x_re = (double)(int32_t)(((d[2 * p] >> sh) & mask) - half_bg);
                    const double x_im = (double)(int32_t)(((d[2 * p + 1] >> sh) & mask) - half_bg);
                    const cplx w = DENSE ? cplx{__hiloint2double((int)twr[4 * p + 1], (int)twr[4 * p]), __hiloint2double((int)twr[4 * p + 3], (int)twr[4 * p + 2])}
                                         : twist[64 * p + t];
                    v[p] = ex_twist(x_re, x_im, w);
                }
                ex_transform<false, DENSE>(v, xb, tw, kc, hi, lo, barid, pr);
                if (DENSE) mbar_wait(&full_bar[stage], phase);
                else
                    while (!mbar_try_wait(&full_bar[stage], phase)) producer_poll(pr);
                const cplx *chunk = bsk_ring + stage * kBskChunkCplx;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    ex_mac(oa[q], v[q], chunk[bsk_slot(0, q, t)]);
                    ex_mac(ob[q], v[q], chunk[bsk_slot(1, q, t)]);
                }
                if (DENSE) ring_release(pr, stage, (uint32_t)((i * 2 + h) * L + l), tid & 31);
                else {
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
                }
                if (++stage == kExStages) { stage = 0; phase ^= 1; }
            }
        }
#pragma unroll 1
        for (int h = 0; h < 2; h++) {     // fft1024 (fft.zig:370-443) of each output + cmux add-back (trgsw.zig:278-281)
            uint32_t *accp = h ? acc_b : acc_a;
            if (h) {
#pragma unroll
                for (int q = 0; q < 8; q++) oa[q] = ob[q];
            }
            ex_transform<true, DENSE>(oa, xb, tw, kc, hi, lo, barid, pr);
            uint32_t twr[32];
            if (DENSE) {
                tmem_ld32(tw.taddr + 64u, twr);
                tmem_wait_ld();
            }
#pragma unroll
            for (int p = 0; p < 8; p++) {
                const cplx w = DENSE ? cplx{__hiloint2double((int)twr[4 * p + 1], (int)twr[4 * p]), __hiloint2double((int)twr[4 * p + 3], (int)twr[4 * p + 2])}
                                     : twist[64 * p + t];
                const cplx r = ex_untwist(oa[p], w);
                double r_re, r_im;
                const uint32_t u_re = ex_round_torus(r.re, &r_re), u_im = ex_round_torus(r.im, &r_im);
                if (MARGIN) margin = fmax(margin, fmax(fabs(r.re - r_re), fabs(r.im - r_im)));
                atomicAdd(&accp[64 * p + t], u_re);          // shared-memory reduction: no load-to-store dependency (see round_accumulate)
                atomicAdd(&accp[64 * p + t + kHalfN], u_im);
            }
        }
        // (no barrier at the end of a step: each half's add-back is followed by a group barrier -- the other half's transform, or the
        // next step's first digit transform -- before the rotated reads that depend on it; tests/test_exchange_protocol.py)
    }
    bar_sync(barid, kGroupThreads);   // the epilogue reads other threads' coefficients

    if (P.out_trlwe) {
        uint32_t *o = P.out_trlwe + ct * (size_t)(2 * kN);
        for (int j = t; j < kN; j += kGroupThreads) { o[j] = acc_a[acc_pos(j)]; o[kN + j] = acc_b[acc_pos(j)]; }
    }
    if (P.out_lv1) {   // sampleExtractIndex(., 0): trlwe.zig:146-162
        uint32_t *o = P.out_lv1 + ct * (size_t)(kN + 1);
        for (int j = t; j <= kN; j += kGroupThreads)
            o[j] = (j == 0) ? acc_a[0] : (j == kN) ? acc_b[0] : 0u - acc_a[acc_pos(kN - j)];
    }
    if (MARGIN && P.margin_bits) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) margin = fmax(margin, __shfl_xor_sync(0xffffffffu, margin, s));
        if ((tid & 31) == 0) atomicMax(P.margin_bits, (unsigned long long)__double_as_longlong(margin));
    }
    if (DENSE) {
        tmem_fence_before_sync();
        bar_sync(15, n_active * kGroupThreads);      // every active warp is done with its tensor-memory window
        if (tid < 32) {
            tmem_fence_after_sync();
            tmem_dealloc(tmem_base, 512);
        }
    }
}

template <int KCT>
cudaError_t launch_rb(const BrArgs &a, const ExactArgs &x, const ExactConsts &kc, bool margin, cudaStream_t s) {
    const size_t smem = ex_fixed_bytes() + (size_t)KCT * ex_group_bytes(a.n, KCT > 4);
    auto k0 = blind_rotate_exact_rb_kernel<KCT, false>;
    auto k1 = blind_rotate_exact_rb_kernel<KCT, true>;
    // gadget shape as compile-time constants at the two widths full waves run at (KCT < 4: generic kernel only)
    constexpr bool kInst = KCT >= 4;
    if (kInst && a.L == 1 && a.bgbit == 22) k0 = blind_rotate_exact_rb_kernel<KCT, false, (kInst ? 1 : 0), (kInst ? 22 : 0)>;        // UINT4 ... UINT8
    else if (kInst && a.L == 1 && a.bgbit == 23) k0 = blind_rotate_exact_rb_kernel<KCT, false, (kInst ? 1 : 0), (kInst ? 23 : 0)>;   // UINT3
    else if (kInst && a.L == 1 && a.bgbit == 18) k0 = blind_rotate_exact_rb_kernel<KCT, false, (kInst ? 1 : 0), (kInst ? 18 : 0)>;   // UINT2
    else if (kInst && a.L == 2 && a.bgbit == 10) k0 = blind_rotate_exact_rb_kernel<KCT, false, (kInst ? 2 : 0), (kInst ? 10 : 0)>;   // UINT1
    // (L = 3 / BGBIT = 6 the same way: 76.1 k against 77.0 k bootstraps/s at the 128-bit set in exact mode -- the unrolled digit loop spills; not instantiated)
    cudaError_t e = cudaFuncSetAttribute(margin ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (a.B + KCT - 1) / KCT;
    if (margin) k1<<<grid, KCT * kGroupThreads, smem, s>>>(a, x.tables, x.bsk_x, x.shared_tab, kc);
    else k0<<<grid, KCT * kGroupThreads, smem, s>>>(a, x.tables, x.bsk_x, x.shared_tab, kc);
    return cudaGetLastError();
}

}  // namespace

// exact_tables: device copy of make_exact_tables() (host_tables.h); bsk_ref: CloudKey.bootstrapping_key as loaded
cudaError_t launch_blind_rotate_exact(const BrArgs &a, const ExactArgs &x, bool track_margin, cudaStream_t s, uint64_t *launches) {
    if (a.B == 0) return cudaSuccess;
    if (!x.legacy && x.bsk_x && x.shared_tab) {
        // host copy of the four pass-A twiddles (kernel parameters): regenerated from the same recurrence as the device tables
        static const ExactConsts kc = [] {
            std::vector<double> tab(6 * 512);
            make_exact_tables(tab.data());
            auto e = [&](int k) { return cplx{tab[2 * 512 + k], tab[3 * 512 + k]}; };
            return ExactConsts{e(2), e(4), e(5), e(6)};
        }();
        const unsigned sms = x.sm_count > 0 ? (unsigned)x.sm_count : 148u;
        int kct = x.kct;
        if (kct < 1 || kct > 6 || kct == 5) {
            // six ciphertexts per CTA (tensor-memory twiddles, profiles/r02_k1x_dense.log) from one full wave of them up, unless
            // margin tracking is on; else the fewest waves of up to four, and below one wave the narrowest CTA that covers the batch
            // (n = 1160, UINT7/8, still fits the 227 KiB of shared memory with 80 bytes to spare; the largest supported n = 1279 does not)
            const bool fits6 = ex_fixed_bytes() + 6 * (size_t)ex_group_bytes(a.n, true) <= 232448;
            if (a.B >= sms * 6 && !track_margin && fits6) kct = 6;
            else {
                kct = 4;
                for (int k = 1; k <= 4; k++)
                    if ((a.B + sms * k - 1) / (sms * k) <= (a.B + sms * 4 - 1) / (sms * 4)) { kct = k; break; }
            }
        }
        if (launches) (*launches)++;
        switch (kct) {
            case 1: return launch_rb<1>(a, x, kc, track_margin, s);
            case 2: return launch_rb<2>(a, x, kc, track_margin, s);
            case 3: return launch_rb<3>(a, x, kc, track_margin, s);
            case 6: return launch_rb<6>(a, x, kc, track_margin, s);
            default: return launch_rb<4>(a, x, kc, track_margin, s);
        }
    }
    const double *exact_tables = x.tables, *bsk_ref = x.bsk_ref;
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

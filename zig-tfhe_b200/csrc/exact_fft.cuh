// exact_fft.cuh -- per-thread building blocks of the EXACT-mode blind rotation (blind_rotate_exact.cu).
//
// Exact mode replays the reference's floating-point DAG operation for operation (same operands, same
// operation order, separate multiplies and adds), so that the accumulator is bit-identical to zig-tfhe's
// on the large-digit parameter sets where the FP64 external product is not an exact integer computation:
//   ifft1024   twist, bit reversal, radix-2 DIT with recurrence twiddles          src/fft.zig:293-366, 582-669
//   fmaInFd1024                                                                   src/trgsw.zig:157-189
//   fft1024    inverse radix-2 DIT, untwist, 1/512, @round                         src/fft.zig:370-443
// What changes is only WHERE each butterfly runs: the nine radix-2 stages are taken three at a time on 8 points
// held in one thread's registers, with the same 64-thread / three-pass / two-exchange
// geometry as the fast kernel (negacyclic_fft.cuh), so the exchanges, the accumulator layout and the key ring
// are shared.  The DAG -- which two values meet in which butterfly with which twiddle -- is untouched.
//
// Index bookkeeping.  Input index k = 64 k2 + 8 k1 + k0 of the folded polynomial sits, after the reference's
// bit reversal (fft.zig:647-669), at position p = brev9(k) = 64 brev3(k0) + 8 brev3(k1) + brev3(k2).  Stage s
// (half = 2^s) pairs positions p and p + half and uses twiddle table entry half - 1 + (p & (half - 1)):
//   pass A = stages 0..2: over k2, thread (k0, k1);         twiddles thread-independent (entries 0..6)
//   pass B = stages 3..5: over k1, thread (k0, l = p & 7);  entries 7 + l, 15 + 8 m + l, 31 + 8 m + l
//   pass C = stages 6..8: over k0, thread q = p & 63;       entries 63 + q, 127 + 64 m + q, 255 + 64 m + q
// A pass takes its 8 registers in input-digit order (register k holds position brev3(k) of the octet) and
// returns them in position order; in unrolled code both are register renaming.  After pass C thread
// (hi, lo) = (l, j1) holds bins p = 64 j0 + 8 lo + hi, j0 = register index -- which is also the role-A
// ownership of the inverse transform, so the pointwise products feed it without an exchange.
//
// Powers of two.  The reference multiplies by 2 after the forward transform, by 0.5 inside the MAC and before
// the inverse transform, and by 1/512 after it.  Scaling by a power of two commutes exactly with every rounded
// multiply and add of the DAG (no value here comes near the subnormal or overflow range), so the device key
// carries the collected factor 2^-10 and those four multiplies disappear; every intermediate is the
// reference's value times an exact power of two and the rounded integer is identical.
// Multiplications by the tabulated twiddle (1, 0) (stage position 0 of pass A) are skipped: x * 1 - y * 0 = x
// exactly, except that the sign of a zero result can differ, which no later operation turns into a different
// non-zero value and the final integer conversion ignores.
#pragma once

#include "negacyclic_fft.cuh"

namespace tfhe_b200 {

constexpr int kExactTabStride = 512;   // make_exact_tables(): twist_re, twist_im, fwd_re, fwd_im, inv_re, inv_im
constexpr int kExactPassATw = 7;       // entries 0..6 of the forward stage table
constexpr int kExactSharedTabCplx = 512;   // twist table as cplx in acc_pos order

// Complex.mul(self = d, other = w) (fft.zig:51-55), then u +- v (fft.zig:605-606).  CONJ: the inverse direction's
// twiddle is the exact conjugate of the forward one (checked on the host), so a - (-b) = a + b reproduces it.
template <bool CONJ>
TFHE_HD void ex_bfly(cplx &u, cplx &d, const cplx w) {
    cplx v;
    if (!CONJ) {
        v.re = sub_(mul_(d.re, w.re), mul_(d.im, w.im));
        v.im = add_(mul_(d.re, w.im), mul_(d.im, w.re));
    } else {
        v.re = add_(mul_(d.re, w.re), mul_(d.im, w.im));
        v.im = sub_(mul_(d.im, w.re), mul_(d.re, w.im));
    }
    const cplx a = cadd(u, v), b = csub(u, v);
    u = a;
    d = b;
}
TFHE_HD void ex_bfly_one(cplx &u, cplx &d) {   // twiddle (1, 0)
    const cplx a = cadd(u, d), b = csub(u, d);
    u = a;
    d = b;
}

// registers in position order <- registers in input-digit order (position j came from register brev3(j))
TFHE_HD void ex_unscramble(cplx (&v)[8]) {
    cplx t;
    t = v[1]; v[1] = v[4]; v[4] = t;
    t = v[3]; v[3] = v[6]; v[6] = t;
}

// three radix-2 DIT stages on one octet.  v: input-digit order in, position order out.
// wa: stage-a twiddle; wb[m], m = position & 1 ... ; wc[m], m = position & 3.
template <bool CONJ>
TFHE_HD void ex_pass(cplx (&v)[8], const cplx wa, const cplx (&wb)[2], const cplx (&wc)[4]) {
    // P(j) = v[brev3(j)]: P0 v0, P1 v4, P2 v2, P3 v6, P4 v1, P5 v5, P6 v3, P7 v7
    ex_bfly<CONJ>(v[0], v[4], wa); ex_bfly<CONJ>(v[2], v[6], wa); ex_bfly<CONJ>(v[1], v[5], wa); ex_bfly<CONJ>(v[3], v[7], wa);
    ex_bfly<CONJ>(v[0], v[2], wb[0]); ex_bfly<CONJ>(v[4], v[6], wb[1]); ex_bfly<CONJ>(v[1], v[3], wb[0]); ex_bfly<CONJ>(v[5], v[7], wb[1]);
    ex_bfly<CONJ>(v[0], v[1], wc[0]); ex_bfly<CONJ>(v[4], v[5], wc[1]); ex_bfly<CONJ>(v[2], v[3], wc[2]); ex_bfly<CONJ>(v[6], v[7], wc[3]);
    ex_unscramble(v);
}
// pass A: table entries 0, 1 and 3 are the recurrence's starting value (1, 0)
template <bool CONJ>
TFHE_HD void ex_pass_a(cplx (&v)[8], const cplx (&w)[kExactPassATw]) {
    ex_bfly_one(v[0], v[4]); ex_bfly_one(v[2], v[6]); ex_bfly_one(v[1], v[5]); ex_bfly_one(v[3], v[7]);
    ex_bfly_one(v[0], v[2]); ex_bfly<CONJ>(v[4], v[6], w[2]); ex_bfly_one(v[1], v[3]); ex_bfly<CONJ>(v[5], v[7], w[2]);
    ex_bfly_one(v[0], v[1]); ex_bfly<CONJ>(v[4], v[5], w[4]); ex_bfly<CONJ>(v[2], v[3], w[5]); ex_bfly<CONJ>(v[6], v[7], w[6]);
    ex_unscramble(v);
}

// per-thread twiddles of pass B (node l = lo of the role-B thread) and pass C (node q = 8 lo + hi of the role-C thread)
struct ExTw {
    cplx wa, wb[2], wc[4];
};
// stage_re / stage_im: forward stage table (entry half - 1 + pos)
TFHE_HD ExTw ex_twiddles_b(const double *stage_re, const double *stage_im, int l) {
    ExTw w;
    w.wa = cplx{stage_re[7 + l], stage_im[7 + l]};
    for (int m = 0; m < 2; m++) w.wb[m] = cplx{stage_re[15 + 8 * m + l], stage_im[15 + 8 * m + l]};
    for (int m = 0; m < 4; m++) w.wc[m] = cplx{stage_re[31 + 8 * m + l], stage_im[31 + 8 * m + l]};
    return w;
}
TFHE_HD ExTw ex_twiddles_c(const double *stage_re, const double *stage_im, int q) {
    ExTw w;
    w.wa = cplx{stage_re[63 + q], stage_im[63 + q]};
    for (int m = 0; m < 2; m++) w.wb[m] = cplx{stage_re[127 + 64 * m + q], stage_im[127 + 64 * m + q]};
    for (int m = 0; m < 4; m++) w.wc[m] = cplx{stage_re[255 + 64 * m + q], stage_im[255 + 64 * m + q]};
    return w;
}

// fold + twist (fft.zig:319-320): z = (x_re c - x_im s, x_re s + x_im c)
TFHE_HD cplx ex_twist(double x_re, double x_im, cplx w) {
    return cplx{sub_(mul_(x_re, w.re), mul_(x_im, w.im)), add_(mul_(x_re, w.im), mul_(x_im, w.re))};
}
// untwist (fft.zig:416-417, the 1/512 lives in the key): (f_re c + f_im s, f_im c - f_re s)
TFHE_HD cplx ex_untwist(cplx f, cplx w) {
    return cplx{add_(mul_(f.re, w.re), mul_(f.im, w.im)), sub_(mul_(f.im, w.re), mul_(f.re, w.im))};
}
// fmaInFd1024 (trgsw.zig:174-181), the x2 / x0.5 pair folded away: acc += (a_re b_re - a_im b_im, a_re b_im + a_im b_re)
TFHE_HD void ex_mac(cplx &acc, cplx a, cplx b) {
    acc.re = add_(acc.re, sub_(mul_(a.re, b.re), mul_(a.im, b.im)));
    acc.im = add_(acc.im, add_(mul_(a.re, b.im), mul_(a.im, b.re)));
}
// @round (half away from zero) -> i64 -> truncating i32 -> u32 (fft.zig:421-424)
TFHE_HD uint32_t ex_round_torus(double t, double *rounded = nullptr) {
#if defined(__CUDA_ARCH__)
    const double r = round(t);
    if (rounded) *rounded = r;
    return (uint32_t)(unsigned long long)__double2ll_rz(r);
#else
    const double r = std::round(t);
    if (rounded) *rounded = r;
    if (r >= 9223372036854775808.0) return 0xFFFFFFFFu;
    if (r < -9223372036854775808.0) return 0u;
    return (uint32_t)(uint64_t)(int64_t)r;
#endif
}

// exact-mode device key: chunk (i, r) = [ab][j0][t] cplx, value = reference bin 64 j0 + 8 (t & 7) + (t >> 3), times 2^-10
TFHE_HD int exact_bin(int j0, int t) { return 64 * j0 + 8 * (t & 7) + (t >> 3); }

}  // namespace tfhe_b200

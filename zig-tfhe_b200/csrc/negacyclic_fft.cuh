// negacyclic_fft.cuh -- per-thread building blocks of the B200 blind-rotation kernel.
//
// Replaces (does not translate) the reference's negacyclic transform pair
//   KlemsaProcessor.ifft1024 / fft1024 / radix2FFT      (src/fft.zig:293-443, 582-669)
// and the pointwise MAC fmaInFd1024                      (src/trgsw.zig:157-189).
//
// Math.  A polynomial p in Z[X]/(X^1024+1) is folded to z_k = p_k + i p_{k+512} (k < 512),
// an element of C[X]/(X^512 - i).  The reference evaluates z at x_j = w^(1-4j), w = exp(i pi/1024)
// (twist by w^k, then a cyclic 512-point radix-2 FFT).  Here the same 512 evaluations are produced
// by three radix-8 passes that split X^M - c into its eight factors X^(M/8) - c_q directly
// (twist merged into the pass twiddles, no separate twist pass, no bit reversal):
//     pass 1: node c = w^512,           r = w^64                       (compile-time constants)
//     pass 2: node q2, c = w^(64+256 q2), r = w^(8+32 q2)              (table tw2)
//     pass 3: node (q2,q1),              r = w^(1+4 q2+32 q1)          (table tw3)
//   each pass: y_p = v_p * r^p (p = 1..7), then an 8-point DFT with kernel exp(+2 pi i p q / 8).
// Leaf (q2,q1,q0) holds z(w^(1+4 q2+32 q1+256 q0)), i.e. the reference's bin j = (-(q2+8 q1+64 q0)) mod 512.
// The pointwise product is order-agnostic, so the bootstrapping key is permuted once at load time
// into this leaf order (and pre-scaled by the exact power of two 1/1024 that collects the
// reference's x2, x2, x0.5, x0.5, 1/512 factors).  The inverse runs the transposed passes with
// conjugate twiddles and no normalisation.
//
// 64 threads own one ciphertext.  Index e = 64 k2 + 8 k1 + k0 of the folded polynomial:
//   role A (passes 1 / 1'):  t = 8 k0 + k1, registers hold k2 = 0..7   (after pass 1: q2 = 0..7)
//   role B (passes 2 / 2'):  t = 8 k0 + q2, registers hold k1 = 0..7   (after pass 2: q1 = 0..7)
//   role C (passes 3 / 3'):  t = 8 q2 + q1, registers hold k0 = 0..7   (after pass 3: q0 = 0..7)
// A<->B is a transpose inside one quarter-warp (exchange X1), B<->C crosses the two warps (X2).
// Both go through padded shared-memory buffers (16-byte slots, row pitches 9 / 73 slots): every
// access is base + compile-time constant, and each quarter-warp touches 8 distinct 16-byte columns,
// so LDS/STS.128 are conflict-free with zero address arithmetic in the loop.
//
// All floating-point contractions are explicit (fma()), and the translation units are compiled
// with --fmad=false / -ffp-contract=off, so the host emulator in tests/ replays the device
// arithmetic bit for bit.
#pragma once

#include <stdint.h>
#if defined(__CUDACC__)
#define TFHE_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define TFHE_HD inline __attribute__((always_inline))
#endif

namespace tfhe_b200 {

struct __attribute__((aligned(16))) cplx {
    double re, im;
};

constexpr int kN = 1024;        // ring degree (every parameter set of the reference)
constexpr int kHalfN = 512;
constexpr int kGroupThreads = 64;  // threads per ciphertext

TFHE_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
TFHE_HD double mul_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
TFHE_HD double add_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
TFHE_HD double sub_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, -b);
#else
    return a - b;
#endif
}

TFHE_HD cplx cadd(cplx a, cplx b) { return cplx{add_(a.re, b.re), add_(a.im, b.im)}; }
TFHE_HD cplx csub(cplx a, cplx b) { return cplx{sub_(a.re, b.re), sub_(a.im, b.im)}; }
// a * w           (2 mul + 2 fma)
TFHE_HD cplx cmul(cplx a, cplx w) {
    return cplx{fma_(-a.im, w.im, mul_(a.re, w.re)), fma_(a.im, w.re, mul_(a.re, w.im))};
}
// a * conj(w)
TFHE_HD cplx cmulc(cplx a, cplx w) {
    return cplx{fma_(a.im, w.im, mul_(a.re, w.re)), fma_(-a.re, w.im, mul_(a.im, w.re))};
}
// acc += a * b    (4 fma)
TFHE_HD void cmac(cplx &acc, cplx a, cplx b) {
    acc.re = fma_(-a.im, b.im, fma_(a.re, b.re, acc.re));
    acc.im = fma_(a.im, b.re, fma_(a.re, b.im, acc.im));
}
// (S*i) * a
template <int S>
TFHE_HD cplx mul_si(cplx a) {
    return (S > 0) ? cplx{-a.im, a.re} : cplx{a.im, -a.re};
}

constexpr double kInvSqrt2 = 0.70710678118654752440084436210485;

// In-place 8-point DFT, natural order in and out:  v[q] <- sum_p v[p] exp(S * 2 pi i p q / 8).
// 52 FP64 pipe operations.
template <int S>
TFHE_HD void dft8(cplx (&v)[8]) {
    const cplx a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
    const cplx c0 = cadd(v[2], v[6]), d0 = mul_si<S>(csub(v[2], v[6]));
    const cplx E0 = cadd(a0, c0), E1 = cadd(b0, d0), E2 = csub(a0, c0), E3 = csub(b0, d0);
    const cplx a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
    const cplx c1 = cadd(v[3], v[7]), d1 = mul_si<S>(csub(v[3], v[7]));
    const cplx O0 = cadd(a1, c1), O1 = cadd(b1, d1), O2 = csub(a1, c1), O3 = csub(b1, d1);
    v[0] = cadd(E0, O0);
    v[4] = csub(E0, O0);
    {   // W^S = (1 + S i)/sqrt2 :  O1 W = ((re - S im) + i (S re + im)) / sqrt2
        const double tr = (S > 0) ? sub_(O1.re, O1.im) : add_(O1.re, O1.im);
        const double ti = (S > 0) ? add_(O1.re, O1.im) : sub_(O1.im, O1.re);
        v[1] = cplx{fma_(kInvSqrt2, tr, E1.re), fma_(kInvSqrt2, ti, E1.im)};
        v[5] = cplx{fma_(-kInvSqrt2, tr, E1.re), fma_(-kInvSqrt2, ti, E1.im)};
    }
    {   // W^2S = S i
        const cplx t = mul_si<S>(O2);
        v[2] = cadd(E2, t);
        v[6] = csub(E2, t);
    }
    {   // W^3S = (-1 + S i)/sqrt2 :  O3 W = ((-re - S im) + i (S re - im)) / sqrt2
        const double tr = (S > 0) ? add_(O3.re, O3.im) : sub_(O3.re, O3.im);   // = -(real part) * sqrt2
        const double ti = (S > 0) ? sub_(O3.re, O3.im) : add_(O3.re, O3.im);   // S>0: +imag*sqrt2 ; S<0: -(imag)*sqrt2
        if (S > 0) {
            v[3] = cplx{fma_(-kInvSqrt2, tr, E3.re), fma_(kInvSqrt2, ti, E3.im)};
            v[7] = cplx{fma_(kInvSqrt2, tr, E3.re), fma_(-kInvSqrt2, ti, E3.im)};
        } else {
            v[3] = cplx{fma_(-kInvSqrt2, tr, E3.re), fma_(-kInvSqrt2, ti, E3.im)};
            v[7] = cplx{fma_(kInvSqrt2, tr, E3.re), fma_(kInvSqrt2, ti, E3.im)};
        }
    }
}

// pass-1 twiddles w^(64 p) = exp(i pi p / 16), p = 0..7 (correctly rounded literals)
TFHE_HD cplx tw1(int p) {
    switch (p) {
        case 1: return cplx{0.98078528040323044912618223613424, 0.19509032201612826784828486847702};
        case 2: return cplx{0.92387953251128675612818318939679, 0.38268343236508977172845998403040};
        case 3: return cplx{0.83146961230254523707878837761791, 0.55557023301960222474283081394853};
        case 4: return cplx{0.70710678118654752440084436210485, 0.70710678118654752440084436210485};
        case 5: return cplx{0.55557023301960222474283081394853, 0.83146961230254523707878837761791};
        case 6: return cplx{0.38268343236508977172845998403040, 0.92387953251128675612818318939679};
        case 7: return cplx{0.19509032201612826784828486847702, 0.98078528040323044912618223613424};
        default: return cplx{1.0, 0.0};
    }
}

// ---- forward passes (twiddle, then DFT8 with + kernel) ----
// (Folding the twiddles into the first butterfly stage -- (x + y w, x - y w) as two chained FMAs per part and 2 x - sum, 6 FP64
// operations instead of 8, 72 instead of 80 per pass, 4.9 % fewer FP64 instructions in the kernel -- was built and measured:
// bit-compatible with the oracle's integers, 102.6 k instead of 103.3 k bootstraps/s at six ciphertexts per CTA and 97.3 k
// instead of 96.3 k at four.  The difference then waits for the sum (chain of 5 instead of 3), and the kernel is bound by
// dependency and shared-memory latency at three warps per scheduler, not by FP64 issue.  profiles/r02_k1_ring.log.)
TFHE_HD void fwd_pass1(cplx (&v)[8]) {
#pragma unroll
    for (int p = 1; p < 8; p++) v[p] = cmul(v[p], tw1(p));
    dft8<+1>(v);
}
// tw points at r^1 for this thread's node; r^p lives at tw[(p-1)*stride]
TFHE_HD void fwd_pass(cplx (&v)[8], const cplx *tw, int stride) {
#pragma unroll
    for (int p = 1; p < 8; p++) v[p] = cmul(v[p], tw[(p - 1) * stride]);
    dft8<+1>(v);
}
// r^1..r^7 from r, r^2, r^4 (4 complex multiplies): trades 16 FP64 operations per pass for 8 registers
// (or 4 shared-memory loads); one extra rounding per derived power, far inside the exactness margin.
TFHE_HD void expand_powers(cplx (&w)[7], cplx r1, cplx r2, cplx r4) {
    w[0] = r1; w[1] = r2; w[2] = cmul(r1, r2); w[3] = r4;
    w[4] = cmul(r1, r4); w[5] = cmul(r2, r4); w[6] = cmul(w[2], r4);
}
// ---- inverse passes (DFT8 with - kernel, then conjugate twiddle) ----
TFHE_HD void inv_pass(cplx (&v)[8], const cplx *tw, int stride) {
    dft8<-1>(v);
#pragma unroll
    for (int p = 1; p < 8; p++) v[p] = cmulc(v[p], tw[(p - 1) * stride]);
}
TFHE_HD void inv_pass1(cplx (&v)[8]) {
    dft8<-1>(v);
#pragma unroll
    for (int p = 1; p < 8; p++) v[p] = cmulc(v[p], tw1(p));
}

// ---- shared-memory exchange slots (units of one cplx = 16 bytes) ----
// X1: role A (k0,k1) element q2  <->  role B (k0,q2) element k1.   Lanes of a quarter-warp differ in
// k1 (writes, stride 1) or q2 (reads, stride 9 = 1 mod 8).
constexpr int kX1Slots = 8 * 72;
TFHE_HD int x1_slot(int k0, int q2, int k1) { return k0 * 72 + q2 * 9 + k1; }
// X2: role B (k0,q2) element q1  <->  role C (q2,q1) element k0.   Writes: lanes differ in q2
// (stride 73 = 1 mod 8); reads: lanes differ in q1 (stride 9).
constexpr int kX2Slots = 7 * 73 + 7 * 9 + 8;
TFHE_HD int x2_slot(int q2, int q1, int k0) { return q2 * 73 + q1 * 9 + k0; }
// X1 laid over an X2 buffer (row pitch 73 like X2's, so that the rows a warp uses for X1 are exactly the rows it alone reads
// in the forward X2 exchange): six ciphertexts per CTA then afford a double-buffered X2 (blind_rotate.cu, Layout::kX1Alias)
TFHE_HD int x1a_slot(int k0, int q2, int k1) { return k0 * 73 + q2 * 9 + k1; }

// ---- accumulator layout in shared memory ----
// Coefficient e of a polynomial lives at acc_pos(e): the two low 3-bit fields of e are swapped, so the
// 16 coefficients e = 64 p + 8 k1 + k0 (p = 0..15) that role-A thread t = 8 k0 + k1 owns sit at
// 64 p + t (conflict-free, constant offsets), and a rotated read (e - atil) is conflict-free as well.
TFHE_HD int acc_pos(int e) { return (e & ~63) | ((e & 7) << 3) | ((e >> 3) & 7); }

// twiddle table layouts (shared memory / global): tw2[p-1][q2], tw3[p-1][t] with t = 8 q2 + q1
TFHE_HD int tw2_index(int p, int q2) { return (p - 1) * 8 + q2; }
TFHE_HD int tw3_index(int p, int t) { return (p - 1) * 64 + t; }
constexpr int kTw2Len = 7 * 8;
constexpr int kTw3Len = 7 * 64;

// device bootstrapping-key layout: chunk (i, r) = [ab][q0][t] cplx, 2*512 cplx = 16 KiB
TFHE_HD int bsk_slot(int ab, int q0, int t) { return ab * 512 + q0 * 64 + t; }
constexpr int kBskChunkCplx = 2 * 512;
constexpr int kBskChunkBytes = kBskChunkCplx * 16;
// reference bin j (fft.zig split layout: re at [j], im at [512+j]) held by leaf (q2,q1,q0)
TFHE_HD int leaf_to_ref_bin(int q2, int q1, int q0) { return (512 - (q2 + 8 * q1 + 64 * q0)) & 511; }

// ---- integer side of one blind-rotation step (trgsw.zig:270-273, 312-321, 442-466, 193-219) ----
// role A: d[2p], d[2p+1] = coefficients e and e+512 of (X^atil * acc - acc) + offset, e = 64 p + 8 k1 + k0,
// atil in [0, 2N].  X^atil * acc at index j is acc[u] for u = (j - atil) mod 2N < N, else -acc[u - N].
// acc is stored in acc_pos order; stepping e by 64 leaves the low 6 position bits unchanged.
TFHE_HD void load_rot_diffs(uint32_t (&d)[16], const uint32_t *acc, int atil, uint32_t offset, int k0, int k1) {
    const int t = 8 * k0 + k1;
    const int u0 = (8 * k1 + k0 - atil) & (2 * kN - 1);
    const int low6 = acc_pos(u0) & 63;
#pragma unroll
    for (int pp = 0; pp < 16; pp++) {          // pp = p (re part) for pp < 8, p + 8 (im part, e + 512) else
        const int u = (u0 + 64 * pp) & (2 * kN - 1);
        const uint32_t v = acc[(u & 0x3C0) | low6];
        const uint32_t rot = (u & kN) ? (0u - v) : v;
        const uint32_t own = acc[64 * pp + t];
        d[(pp & 7) * 2 + (pp >> 3)] = rot - own + offset;
    }
}
// gadget digit of level l (shift sh = 32 - (l+1)*bgbit) as signed integer -> double
// exact int32 -> double.  I2F.F64.S32 runs on the conversion unit, in parallel with the FP64 pipe.  (Measured: assembling
// the double from its bit pattern + one DADD instead moves 96 operations per warp and CMUX step onto the FP64 pipe,
// the busier of the two: 88.5 k instead of 90.9 k bootstraps/s.)
TFHE_HD double int_to_double(int32_t x) { return (double)x; }
TFHE_HD void digits_to_cplx(cplx (&v)[8], const uint32_t (&d)[16], int sh, uint32_t mask, uint32_t half_bg) {
#pragma unroll
    for (int p = 0; p < 8; p++) {
        v[p].re = int_to_double((int32_t)(((d[2 * p] >> sh) & mask) - half_bg));
        v[p].im = int_to_double((int32_t)(((d[2 * p + 1] >> sh) & mask) - half_bg));
    }
}

// round-to-nearest of an (almost) integer double, reduced mod 2^32 (fft.zig:421-424).
// magic = 1.5 * 2^52 trick: valid for |x| < 2^51 (all L=3/BGBIT=6 sets: |x| < 2^45).
TFHE_HD uint32_t round_torus_magic(double x) {
    const double t = add_(x, 6755399441055744.0);
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(t);
#else
    uint64_t bits;
    __builtin_memcpy(&bits, &t, 8);
    return (uint32_t)bits;
#endif
}
// general version (large-digit sets): round half to even, saturating to i64 like F2I.S64
TFHE_HD uint32_t round_torus_wide(double x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)(unsigned long long)__double2ll_rn(x);
#else
    const double r = std::nearbyint(x);
    if (r >= 9223372036854775808.0) return 0xFFFFFFFFu;
    if (r < -9223372036854775808.0) return 0u;
    return (uint32_t)(uint64_t)(int64_t)r;
#endif
}

}  // namespace tfhe_b200

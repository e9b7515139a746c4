// host_tables.h -- host-side generation of the pass-2 / pass-3 twiddle tables of negacyclic_fft.cuh.
// (The reference builds its twists with @cos/@sin at plan creation, src/fft.zig:98-106; here the
// tables are computed in long double with octant symmetry and rounded once to double.)
#pragma once
#include <cmath>

#include "exact_fft.cuh"
#include "negacyclic_fft.cuh"

namespace tfhe_b200 {

// exp(i pi k / 1024), k taken mod 2048, exactly symmetric across octants
inline cplx unit_root_2048(int k) {
    k &= 2047;
    const int quad = k >> 9, r = k & 511;
    long double c, s;
    const long double pi = 3.14159265358979323846264338327950288419716939937510L;
    if (r == 0) { c = 1.0L; s = 0.0L; }
    else if (r <= 256) { c = cosl(pi * r / 1024.0L); s = sinl(pi * r / 1024.0L); }
    else { c = sinl(pi * (512 - r) / 1024.0L); s = cosl(pi * (512 - r) / 1024.0L); }
    if (r == 256) { c = s = 0.70710678118654752440084436210485L; }
    double cd = (double)c, sd = (double)s;
    switch (quad) {
        case 0: return cplx{cd, sd};
        case 1: return cplx{-sd, cd};
        case 2: return cplx{-cd, -sd};
        default: return cplx{sd, -cd};
    }
}

// tw2[tw2_index(p,q2)] = w^((8+32 q2) p), tw3[tw3_index(p,t)] = w^((1+4 q2+32 q1) p), t = 8 q2 + q1
inline void make_twiddle_tables(cplx *tw2, cplx *tw3) {
    for (int p = 1; p < 8; p++) {
        for (int q2 = 0; q2 < 8; q2++) tw2[tw2_index(p, q2)] = unit_root_2048((8 + 32 * q2) * p);
        for (int t = 0; t < 64; t++) {
            const int q2 = t >> 3, q1 = t & 7;
            tw3[tw3_index(p, t)] = unit_root_2048((1 + 4 * q2 + 32 * q1) * p);
        }
    }
}

// Exact-mode tables: the very values the reference's plan and radix-2 loop use (src/fft.zig:98-106 twists,
// :589-611 stage twiddles from the serial recurrence w <- w * w_len), 6 x 512 doubles:
//   [0] twist_re  [1] twist_im  [2] fwd_re  [3] fwd_im  [4] inv_re  [5] inv_im ; stage s at [2^s - 1, 2^(s+1) - 1).
// Compiled with -ffp-contract=off so the recurrence is evaluated with separate multiplies and adds.
inline void make_exact_tables(double *out) {
    const double pi = 3.14159265358979323846;
    const double twist_unit = pi / 1024.0;                         // fft.zig:101
    for (int i = 0; i < 512; i++) {
        const double angle = (double)i * twist_unit;               // fft.zig:103
        out[i] = std::cos(angle);
        out[512 + i] = std::sin(angle);
    }
    for (int inv = 0; inv < 2; inv++) {
        double *tr = out + (2 + 2 * inv) * 512, *ti = tr + 512;
        tr[511] = ti[511] = 0.0;
        for (int len = 2; len <= 512; len *= 2) {
            const double angle = inv ? 2.0 * pi / (double)len : -2.0 * pi / (double)len;   // fft.zig:591
            const double wl_re = std::cos(angle), wl_im = std::sin(angle);                   // fft.zig:592-593
            double w_re = 1.0, w_im = 0.0;
            for (int j = 0; j < len / 2; j++) {
                tr[len / 2 - 1 + j] = w_re;
                ti[len / 2 - 1 + j] = w_im;
                const double temp = w_re * wl_re - w_im * wl_im;                             // fft.zig:609
                w_im = w_re * wl_im + w_im * wl_re;                                          // fft.zig:610
                w_re = temp;
            }
        }
    }
}

// true when every inverse stage twiddle is the exact conjugate of the forward one (cos even, sin odd in this libm, and
// the recurrence then conjugates exactly): the register-blocked exact kernel keeps one set of twiddles for both directions
inline bool exact_tables_conjugate(const double *tab) {
    const double *fr = tab + 2 * 512, *fi = tab + 3 * 512, *ir = tab + 4 * 512, *ii = tab + 5 * 512;
    for (int k = 0; k < 511; k++)
        if (fr[k] != ir[k] || fi[k] != -ii[k]) return false;
    // the skipped multiplications of pass A rely on these entries being exactly (1, 0)
    for (int k : {0, 1, 3})
        if (fr[k] != 1.0 || fi[k] != 0.0) return false;
    return true;
}

// shared-memory image of the exact kernel: twist[512] as cplx in acc_pos order + the 7 thread-independent pass-A twiddles
inline void make_exact_shared_tables(const double *tab, cplx *twist, cplx *twa) {
    for (int k = 0; k < 512; k++) twist[acc_pos(k)] = cplx{tab[k], tab[512 + k]};
    for (int k = 0; k < kExactPassATw; k++) twa[k] = cplx{tab[2 * 512 + k], tab[3 * 512 + k]};
}

}  // namespace tfhe_b200

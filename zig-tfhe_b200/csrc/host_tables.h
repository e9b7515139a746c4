// host_tables.h -- host-side generation of the pass-2 / pass-3 twiddle tables of negacyclic_fft.cuh.
// (The reference builds its twists with @cos/@sin at plan creation, src/fft.zig:98-106; here the
// tables are computed in long double with octant symmetry and rounded once to double.)
#pragma once
#include <cmath>

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

}  // namespace tfhe_b200

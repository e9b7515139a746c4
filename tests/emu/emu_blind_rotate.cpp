// emu_blind_rotate.cpp -- HOST EMULATOR of the fast-mode blind-rotation kernel (test infrastructure).
//
// Replays, thread by thread and barrier phase by barrier phase, exactly what one 64-thread group of
// zig-tfhe_b200/csrc/blind_rotate.cu does, using the very same __host__ __device__ building blocks
// (negacyclic_fft.cuh).  Because every FP contraction in those blocks is an explicit fma and this
// file is compiled with -ffp-contract=off, the emulator's output is bit-identical to the GPU's.
// It exists so the algorithm (radix-8 split, swizzles, key permutation, rounding) can be validated
// against the oracle on a machine with no GPU.  It is NOT part of the product and is never loaded
// by it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../zig-tfhe_b200/csrc/exact_fft.cuh"
#include "../../zig-tfhe_b200/csrc/host_tables.h"
#include "../../zig-tfhe_b200/csrc/negacyclic_fft.cuh"

using namespace tfhe_b200;

namespace {

struct Group {
    uint32_t acc_a[kN], acc_b[kN];
    cplx x1[kX1Slots], x2[kX2Slots];
    cplx tw2[kTw2Len], tw3[kTw3Len];
    Group() { make_twiddle_tables(tw2, tw3); }
};

// forward transform of all 64 threads, v[t][8]: role A in -> role C out (leaf order)
void fwd_all(Group &G, cplx (*v)[8]) {
    for (int t = 0; t < 64; t++) {  // pass 1 + X1 write (role A: k0 = hi, k1 = lo)
        const int hi = t >> 3, lo = t & 7;
        fwd_pass1(v[t]);
        for (int q2 = 0; q2 < 8; q2++) G.x1[x1_slot(hi, q2, lo)] = v[t][q2];
    }
    for (int t = 0; t < 64; t++) {  // X1 read (role B: k0 = hi, q2 = lo) + pass 2
        const int hi = t >> 3, lo = t & 7;
        for (int k1 = 0; k1 < 8; k1++) v[t][k1] = G.x1[x1_slot(hi, lo, k1)];
        fwd_pass(v[t], G.tw2 + lo, 8);
    }
    for (int t = 0; t < 64; t++) {  // X2 write
        const int hi = t >> 3, lo = t & 7;
        for (int q1 = 0; q1 < 8; q1++) G.x2[x2_slot(lo, q1, hi)] = v[t][q1];
    }
    for (int t = 0; t < 64; t++) {  // X2 read (role C: q2 = hi, q1 = lo) + pass 3
        const int hi = t >> 3, lo = t & 7;
        for (int k0 = 0; k0 < 8; k0++) v[t][k0] = G.x2[x2_slot(hi, lo, k0)];
        fwd_pass(v[t], G.tw3 + t, 64);
    }
}

// inverse transform: role C in (leaf order) -> role A out (v[t][p] = c_e, e = 64 p + 8 lo + hi)
void inv_all(Group &G, cplx (*v)[8]) {
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        inv_pass(v[t], G.tw3 + t, 64);
        for (int k0 = 0; k0 < 8; k0++) G.x2[x2_slot(hi, lo, k0)] = v[t][k0];
    }
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        for (int q1 = 0; q1 < 8; q1++) v[t][q1] = G.x2[x2_slot(lo, q1, hi)];
        inv_pass(v[t], G.tw2 + lo, 8);
        for (int k1 = 0; k1 < 8; k1++) G.x1[x1_slot(hi, lo, k1)] = v[t][k1];
    }
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        for (int q2 = 0; q2 < 8; q2++) v[t][q2] = G.x1[x1_slot(hi, q2, lo)];
        inv_pass1(v[t]);
    }
}

}  // namespace

extern "C" {

// reference CloudKey.bootstrapping_key [n][2L][2][N] f64  ->  device layout [n*2L][ab][q0][t] cplx / 1024
void emu_permute_bsk(const double *ref, int n, int L, double *out) {
    cplx *o = reinterpret_cast<cplx *>(out);
    for (size_t c = 0; c < (size_t)n * 2 * L; c++)
        for (int ab = 0; ab < 2; ab++) {
            const double *src = ref + (c * 2 + ab) * kN;
            for (int q0 = 0; q0 < 8; q0++)
                for (int t = 0; t < 64; t++) {
                    const int j = leaf_to_ref_bin(t >> 3, t & 7, q0);
                    o[c * kBskChunkCplx + bsk_slot(ab, q0, t)] = cplx{src[j] * (1.0 / 1024.0), src[kHalfN + j] * (1.0 / 1024.0)};
                }
        }
}

// forward transform of one polynomial given as signed integers; out[512] cplx in slot order q0*64+t
void emu_forward(const int32_t *poly, double *out) {
    static Group G;
    cplx v[64][8];
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        for (int p = 0; p < 8; p++) {
            const int e = 64 * p + 8 * lo + hi;
            v[t][p] = cplx{(double)poly[e], (double)poly[e + kHalfN]};
        }
    }
    fwd_all(G, v);
    cplx *o = reinterpret_cast<cplx *>(out);
    for (int t = 0; t < 64; t++)
        for (int q0 = 0; q0 < 8; q0++) o[q0 * 64 + t] = v[t][q0];
}

// inverse (unnormalised) of a spectrum in slot order; out[1024] doubles (coefficient k, k+512)
void emu_inverse(const double *spec, double *out) {
    static Group G;
    const cplx *s = reinterpret_cast<const cplx *>(spec);
    cplx v[64][8];
    for (int t = 0; t < 64; t++)
        for (int q0 = 0; q0 < 8; q0++) v[t][q0] = s[q0 * 64 + t];
    inv_all(G, v);
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        for (int p = 0; p < 8; p++) {
            const int e = 64 * p + 8 * lo + hi;
            out[e] = v[t][p].re;
            out[e + kHalfN] = v[t][p].im;
        }
    }
}

// full blind rotation of one ciphertext `lin` (already the gate's linear combination), fast mode.
// bskp: device-layout key from emu_permute_bsk.  testvec NULL -> (0, 2^29).  trace optional [n][2][N].
void emu_blind_rotate(int n, int L, int bgbit, uint32_t offset, const uint32_t *lin, const double *bskp,
                      const uint32_t *testvec, int wide_round, uint32_t *out_trlwe, uint32_t *trace, double *margin_out) {
    static Group G;
    const cplx *bsk = reinterpret_cast<const cplx *>(bskp);
    const uint32_t mask = (1u << bgbit) - 1u, half_bg = 1u << (bgbit - 1);
    std::vector<int> atil(n + 1);
    for (int i = 0; i <= n; i++) {
        const uint32_t m = (uint32_t)(((uint64_t)lin[i] + (1u << 20)) >> 21);
        atil[i] = (i == n) ? (2 * kN - (int)m) : (int)m;
    }
    const int btil = atil[n];
    for (int j = 0; j < kN; j++) {
        const int u = (j - btil) & (2 * kN - 1);
        const uint32_t va = testvec ? testvec[u & (kN - 1)] : 0u;
        const uint32_t vb = testvec ? testvec[kN + (u & (kN - 1))] : 0x20000000u;
        G.acc_a[acc_pos(j)] = (u & kN) ? 0u - va : va;   // accumulators live in acc_pos order, like the kernel's
        G.acc_b[acc_pos(j)] = (u & kN) ? 0u - vb : vb;
    }
    double margin = 0.0;
    static cplx v[64][8], oa[64][8], ob[64][8];
    for (int i = 0; i < n; i++) {
        std::memset(oa, 0, sizeof(oa));
        std::memset(ob, 0, sizeof(ob));
        for (int h = 0; h < 2; h++) {
            const uint32_t *accp = h ? G.acc_b : G.acc_a;
            static uint32_t d[64][16];
            for (int t = 0; t < 64; t++) load_rot_diffs(d[t], accp, atil[i], offset, t >> 3, t & 7);
            for (int l = 0; l < L; l++) {
                for (int t = 0; t < 64; t++) digits_to_cplx(v[t], d[t], 32 - (l + 1) * bgbit, mask, half_bg);
                fwd_all(G, v);
                const cplx *chunk = bsk + ((size_t)i * 2 * L + h * L + l) * kBskChunkCplx;
                for (int t = 0; t < 64; t++)
                    for (int q0 = 0; q0 < 8; q0++) {
                        cmac(oa[t][q0], v[t][q0], chunk[bsk_slot(0, q0, t)]);
                        cmac(ob[t][q0], v[t][q0], chunk[bsk_slot(1, q0, t)]);
                    }
            }
        }
        for (int h = 0; h < 2; h++) {
            cplx(*o)[8] = h ? ob : oa;
            uint32_t *accp = h ? G.acc_b : G.acc_a;
            inv_all(G, o);
            for (int t = 0; t < 64; t++) {
                for (int p = 0; p < 8; p++) {
                    const int e = 64 * p + t;   // acc_pos of coefficient 64 p + 8 lo + hi
                    const double xr = o[t][p].re, xi = o[t][p].im;
                    margin = std::fmax(margin, std::fabs(xr - std::nearbyint(xr)));
                    margin = std::fmax(margin, std::fabs(xi - std::nearbyint(xi)));
                    accp[e] += wide_round ? round_torus_wide(xr) : round_torus_magic(xr);
                    accp[e + kHalfN] += wide_round ? round_torus_wide(xi) : round_torus_magic(xi);
                }
            }
        }
        if (trace)
            for (int j = 0; j < kN; j++) {
                trace[(size_t)i * 2 * kN + j] = G.acc_a[acc_pos(j)];
                trace[(size_t)i * 2 * kN + kN + j] = G.acc_b[acc_pos(j)];
            }
    }
    for (int j = 0; j < kN; j++) {
        out_trlwe[j] = G.acc_a[acc_pos(j)];
        out_trlwe[kN + j] = G.acc_b[acc_pos(j)];
    }
    if (margin_out) *margin_out = margin;
}


// ---- exact mode (blind_rotate_exact.cu, exact_fft.cuh): the reference DAG, three radix-2 stages per register pass ----

// reference CloudKey.bootstrapping_key -> exact-mode device layout [n*2L][ab][j0][t] cplx, times 2^-10
void emu_permute_bsk_exact(const double *ref, int n, int L, double *out) {
    cplx *o = reinterpret_cast<cplx *>(out);
    for (size_t c = 0; c < (size_t)n * 2 * L; c++)
        for (int ab = 0; ab < 2; ab++) {
            const double *src = ref + (c * 2 + ab) * kN;
            for (int j0 = 0; j0 < 8; j0++)
                for (int t = 0; t < 64; t++) {
                    const int j = exact_bin(j0, t);
                    o[c * kBskChunkCplx + bsk_slot(ab, j0, t)] = cplx{src[j] * (1.0 / 1024.0), src[kHalfN + j] * (1.0 / 1024.0)};
                }
        }
}

int emu_exact_tables_conjugate() {
    std::vector<double> tab(6 * 512);
    make_exact_tables(tab.data());
    return exact_tables_conjugate(tab.data()) ? 1 : 0;
}

}  // extern "C"

namespace {
struct ExGroup {
    uint32_t acc_a[kN], acc_b[kN];
    cplx x1[kX1Slots], x2[kX2Slots];
    cplx twist[512];            // acc_pos order
    cplx twa[kExactPassATw];
    ExTw twb[64], twc[64];      // per thread
    ExGroup() {
        std::vector<double> tab(6 * 512);
        make_exact_tables(tab.data());
        make_exact_shared_tables(tab.data(), twist, twa);
        for (int t = 0; t < 64; t++) {
            twb[t] = ex_twiddles_b(tab.data() + 2 * 512, tab.data() + 3 * 512, t & 7);
            twc[t] = ex_twiddles_c(tab.data() + 2 * 512, tab.data() + 3 * 512, 8 * (t & 7) + (t >> 3));
        }
    }
};

// role A registers in (input-digit order) -> role C registers out (position order = bins 64 j0 + 8 lo + hi)
template <bool CONJ>
void ex_transform_all(ExGroup &G, cplx (*v)[8]) {
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        ex_pass_a<CONJ>(v[t], G.twa);
        for (int q = 0; q < 8; q++) G.x1[x1_slot(hi, q, lo)] = v[t][q];
    }
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        for (int q = 0; q < 8; q++) v[t][q] = G.x1[x1_slot(hi, lo, q)];
        ex_pass<CONJ>(v[t], G.twb[t].wa, G.twb[t].wb, G.twb[t].wc);
    }
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        for (int q = 0; q < 8; q++) G.x2[x2_slot(lo, q, hi)] = v[t][q];
    }
    for (int t = 0; t < 64; t++) {
        const int hi = t >> 3, lo = t & 7;
        for (int q = 0; q < 8; q++) v[t][q] = G.x2[x2_slot(hi, lo, q)];
        ex_pass<CONJ>(v[t], G.twc[t].wa, G.twc[t].wb, G.twc[t].wc);
    }
}
}  // namespace

extern "C" {

// forward transform alone: poly[1024] signed -> out[1024] doubles in the REFERENCE layout (re | im), times 1/2
void emu_exact_forward(const int32_t *poly, double *out) {
    static ExGroup G;
    static cplx v[64][8];
    for (int t = 0; t < 64; t++)
        for (int p = 0; p < 8; p++) {
            const int e = 64 * p + 8 * (t & 7) + (t >> 3);
            v[t][p] = ex_twist((double)poly[e], (double)poly[e + kHalfN], G.twist[64 * p + t]);
        }
    ex_transform_all<false>(G, v);
    for (int t = 0; t < 64; t++)
        for (int j0 = 0; j0 < 8; j0++) {
            out[exact_bin(j0, t)] = v[t][j0].re;
            out[kHalfN + exact_bin(j0, t)] = v[t][j0].im;
        }
}

void emu_blind_rotate_exact(int n, int L, int bgbit, uint32_t offset, const uint32_t *lin, const double *bskx,
                            const uint32_t *testvec, uint32_t *out_trlwe, uint32_t *trace) {
    static ExGroup G;
    const cplx *bsk = reinterpret_cast<const cplx *>(bskx);
    const uint32_t mask = (1u << bgbit) - 1u, half_bg = 1u << (bgbit - 1);
    std::vector<int> atil(n + 1);
    for (int i = 0; i <= n; i++) {
        const uint32_t m = (uint32_t)(((uint64_t)lin[i] + (1u << 20)) >> 21);
        atil[i] = (i == n) ? (2 * kN - (int)m) : (int)m;
    }
    const int btil = atil[n];
    for (int j = 0; j < kN; j++) {
        const int u = (j - btil) & (2 * kN - 1);
        const uint32_t va = testvec ? testvec[u & (kN - 1)] : 0u;
        const uint32_t vb = testvec ? testvec[kN + (u & (kN - 1))] : 0x20000000u;
        G.acc_a[acc_pos(j)] = (u & kN) ? 0u - va : va;
        G.acc_b[acc_pos(j)] = (u & kN) ? 0u - vb : vb;
    }
    static cplx v[64][8], oa[64][8], ob[64][8];
    for (int i = 0; i < n; i++) {
        std::memset(oa, 0, sizeof(oa));
        std::memset(ob, 0, sizeof(ob));
        for (int h = 0; h < 2; h++) {
            const uint32_t *accp = h ? G.acc_b : G.acc_a;
            static uint32_t d[64][16];
            for (int t = 0; t < 64; t++) load_rot_diffs(d[t], accp, atil[i], offset, t >> 3, t & 7);
            for (int l = 0; l < L; l++) {
                const int sh = 32 - (l + 1) * bgbit;
                for (int t = 0; t < 64; t++)
                    for (int p = 0; p < 8; p++) {
                        const double x_re = (double)(int32_t)(((d[t][2 * p] >> sh) & mask) - half_bg);
                        const double x_im = (double)(int32_t)(((d[t][2 * p + 1] >> sh) & mask) - half_bg);
                        v[t][p] = ex_twist(x_re, x_im, G.twist[64 * p + t]);
                    }
                ex_transform_all<false>(G, v);
                const cplx *chunk = bsk + ((size_t)i * 2 * L + h * L + l) * kBskChunkCplx;
                for (int t = 0; t < 64; t++)
                    for (int q = 0; q < 8; q++) {
                        ex_mac(oa[t][q], v[t][q], chunk[bsk_slot(0, q, t)]);
                        ex_mac(ob[t][q], v[t][q], chunk[bsk_slot(1, q, t)]);
                    }
            }
        }
        for (int h = 0; h < 2; h++) {
            cplx(*o)[8] = h ? ob : oa;
            uint32_t *accp = h ? G.acc_b : G.acc_a;
            ex_transform_all<true>(G, o);
            for (int t = 0; t < 64; t++)
                for (int p = 0; p < 8; p++) {
                    const cplx r = ex_untwist(o[t][p], G.twist[64 * p + t]);
                    accp[64 * p + t] += ex_round_torus(r.re);
                    accp[64 * p + t + kHalfN] += ex_round_torus(r.im);
                }
        }
        if (trace)
            for (int j = 0; j < kN; j++) {
                trace[(size_t)i * 2 * kN + j] = G.acc_a[acc_pos(j)];
                trace[(size_t)i * 2 * kN + kN + j] = G.acc_b[acc_pos(j)];
            }
    }
    for (int j = 0; j < kN; j++) {
        out_trlwe[j] = G.acc_a[acc_pos(j)];
        out_trlwe[kN + j] = G.acc_b[acc_pos(j)];
    }
}

}  // extern "C"

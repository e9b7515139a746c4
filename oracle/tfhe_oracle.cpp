// tfhe_oracle.cpp -- CPU ORACLE (test infrastructure only; see tfhe_oracle.h header comment).
//
// Operation-for-operation restatement of zig-tfhe's gate-bootstrapping path.  All
// citations are file:line under /root/reference/.  Compile with -ffp-contract=off.
//
// Deliberate, value-preserving deviations from the reference text:
//  * radix-2 stage twiddles: the reference re-runs the serial recurrence w <- w*w_len for
//    every block of every call (fft.zig:596-611); the sequence depends only on (len, j), so
//    it is tabulated once with the very same recurrence.  Values are bit-identical.
//  * cos/sin come from glibc, the reference uses Zig's @cos/@sin (fft.zig:104-105,592-593);
//    they may differ in the last ulp.
//  * PRNG: the reference seeds Xoshiro256++ from a clock (utils.zig:16-22, non-reproducible);
//    here every encryption draws from a seeded xoshiro256++ so keys/ciphertexts are replayable.
#include "tfhe_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

constexpr int N1 = 1024;   // trgsw_lv1.N, params.zig:85 ... 365 (every set)
constexpr int N2 = 512;
constexpr int NBIT = 10;
constexpr int TORUS_SIZE = 32;  // params.zig:30

// ------------------------------------------------------------------ PRNG (seeded stand-in)
struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto &v : s) v = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {  // xoshiro256++
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    uint32_t u32() { return (uint32_t)(next() >> 32); }
    bool boolean() { return (next() >> 63) != 0; }
    double f64() { return (double)(next() >> 11) * 0x1.0p-53; }  // [0,1)
};

// utils.zig:50-82 Box-Muller with spare
struct NormalDist {
    double mean, stddev; bool has_spare = false; double spare = 0.0;
    NormalDist(double m, double s) : mean(m), stddev(s) {}
    double next(Rng &rng) {
        if (has_spare) { has_spare = false; return spare * stddev + mean; }
        const double u1 = rng.f64();
        const double u2 = rng.f64();
        const double mag = stddev * std::sqrt(-2.0 * std::log(u1));
        const double z0 = mag * std::cos(2.0 * M_PI * u2);
        const double z1 = mag * std::sin(2.0 * M_PI * u2);
        has_spare = true; spare = z1;
        return z0 + mean;
    }
};

// utils.zig:28-33.  Zig lowers @mod on runtime floats as
//   a = fmod(x,y); lhs<0 ? fmod(a+y, y) : a
// (so a tiny negative input maps to 0, not to 1-ulp).
inline uint32_t f64_to_torus(double d) {
    double a = std::fmod(d, 1.0);
    double normalized = (d < 0.0) ? std::fmod(a + 1.0, 1.0) : a;
    double torus = normalized * 4294967296.0;
    double clamped = std::max(0.0, std::min(torus, 4294967295.0));
    return (uint32_t)clamped;  // @intFromFloat truncates
}

// utils.zig:85-104
inline uint32_t gaussian_torus(uint32_t mu, NormalDist &nd, Rng &rng) { return f64_to_torus(nd.next(rng)) + mu; }
inline uint32_t gaussian_f64(double mu, NormalDist &nd, Rng &rng) { return gaussian_torus(f64_to_torus(mu), nd, rng); }

// ------------------------------------------------------------------ FFT plan (fft.zig:79-129)
struct Plan {
    double tw_re[N2], tw_im[N2];                 // twisties, fft.zig:98-106
    // stage twiddle sequences from the reference recurrence (fft.zig:589-611), fwd and inv
    std::vector<double> st_re[2][10], st_im[2][10];
    int brev[N2];
    Plan() {
        const double twist_unit = M_PI / (double)N1;           // fft.zig:101
        for (int i = 0; i < N2; i++) {
            const double angle = (double)i * twist_unit;        // fft.zig:103
            tw_re[i] = std::cos(angle); tw_im[i] = std::sin(angle);
        }
        for (int inv = 0; inv < 2; inv++) {
            int s = 0;
            for (int len = 2; len <= N2; len *= 2, s++) {
                const double angle = inv ? 2.0 * M_PI / (double)len : -2.0 * M_PI / (double)len;  // fft.zig:591
                const double wl_re = std::cos(angle), wl_im = std::sin(angle);                      // fft.zig:592-593
                double w_re = 1.0, w_im = 0.0;                                                      // fft.zig:597-598
                st_re[inv][s].resize(len / 2); st_im[inv][s].resize(len / 2);
                for (int j = 0; j < len / 2; j++) {
                    st_re[inv][s][j] = w_re; st_im[inv][s][j] = w_im;
                    const double temp = w_re * wl_re - w_im * wl_im;                                // fft.zig:609
                    w_im = w_re * wl_im + w_im * wl_re;                                             // fft.zig:610
                    w_re = temp;                                                                    // fft.zig:611
                }
            }
        }
        // fft.zig:647-669 (incrementing-j loop) == plain 9-bit reversal
        int j = 0;
        for (int i = 0; i < N2; i++) {
            brev[i] = j;
            int mask = N2 >> 1;
            while (mask > 0 && (j & mask) != 0) { j ^= mask; mask >>= 1; }
            j ^= mask;
        }
    }
};
const Plan &plan() { static Plan p; return p; }

// fft.zig:582-619 radix-2 DIT on split re/im arrays of length 512
inline void radix2_512(double *re, double *im, bool inverse) {
    const Plan &P = plan();
    for (int i = 0; i < N2; i++) {               // fft.zig:647-669
        int j = P.brev[i];
        if (j > i) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    int s = 0;
    for (int len = 2; len <= N2; len *= 2, s++) {
        const int half = len / 2;
        const double *wr = P.st_re[inverse ? 1 : 0][s].data();
        const double *wi = P.st_im[inverse ? 1 : 0][s].data();
        for (int i = 0; i < N2; i += len) {
            for (int j = 0; j < half; j++) {
                const double u_re = re[i + j], u_im = im[i + j];
                const double d_re = re[i + j + half], d_im = im[i + j + half];
                // Complex.mul, fft.zig:51-55 (self = data, other = w)
                const double v_re = d_re * wr[j] - d_im * wi[j];
                const double v_im = d_re * wi[j] + d_im * wr[j];
                re[i + j] = u_re + v_re; im[i + j] = u_im + v_im;               // fft.zig:605
                re[i + j + half] = u_re - v_re; im[i + j + half] = u_im - v_im; // fft.zig:606
            }
        }
    }
}

// fft.zig:293-366
inline void ifft1024(const uint32_t *in, double *out) {
    const Plan &P = plan();
    double re[N2], im[N2];
    for (int i = 0; i < N2; i++) {
        const double in_re = (double)(int32_t)in[i];
        const double in_im = (double)(int32_t)in[i + N2];
        re[i] = in_re * P.tw_re[i] - in_im * P.tw_im[i];   // fft.zig:319
        im[i] = in_re * P.tw_im[i] + in_im * P.tw_re[i];   // fft.zig:320
    }
    radix2_512(re, im, false);
    for (int i = 0; i < N2; i++) { out[i] = re[i] * 2.0; out[i + N2] = im[i] * 2.0; }  // fft.zig:356-357
}

// fft.zig:370-443
inline void fft1024(const double *in, uint32_t *out, double *max_frac) {
    const Plan &P = plan();
    double re[N2], im[N2];
    for (int i = 0; i < N2; i++) { re[i] = in[i] * 0.5; im[i] = in[i + N2] * 0.5; }     // fft.zig:380
    radix2_512(re, im, true);
    const double normalization = 1.0 / (double)N2;                                       // fft.zig:392
    for (int i = 0; i < N2; i++) {
        const double tmp_re = (re[i] * P.tw_re[i] + im[i] * P.tw_im[i]) * normalization; // fft.zig:416
        const double tmp_im = (im[i] * P.tw_re[i] - re[i] * P.tw_im[i]) * normalization; // fft.zig:417
        const double r_re = std::round(tmp_re), r_im = std::round(tmp_im);               // @round: half away from zero
        if (max_frac) {
            *max_frac = std::max(*max_frac, std::fabs(tmp_re - r_re));
            *max_frac = std::max(*max_frac, std::fabs(tmp_im - r_im));
        }
        out[i] = (uint32_t)(int32_t)(int64_t)r_re;          // i64 -> truncate i32 -> bitcast, fft.zig:421-424
        out[i + N2] = (uint32_t)(int32_t)(int64_t)r_im;
    }
}

// fft.zig:458-492
inline void poly_mul_fft(const uint32_t *a, const uint32_t *b, uint32_t *out) {
    double af[N1], bf[N1], rf[N1];
    ifft1024(a, af); ifft1024(b, bf);
    for (int i = 0; i < N2; i++) {
        const double ar = af[i], ai = af[i + N2], br = bf[i], bi = bf[i + N2];
        rf[i] = (ar * br - ai * bi) * 0.5;        // fft.zig:478
        rf[i + N2] = (ar * bi + ai * br) * 0.5;   // fft.zig:479
    }
    fft1024(rf, out, nullptr);
}

// trgsw.zig:157-189
inline void fma_in_fd_1024(double *res, const double *a, const double *b) {
    for (int i = 0; i < N2; i++) {
        res[i] = res[i] + (a[i] * b[i] - a[i + N2] * b[i + N2]) * 0.5;
        res[i + N2] = res[i + N2] + (a[i] * b[i + N2] + a[i + N2] * b[i]) * 0.5;
    }
}

// trgsw.zig:193-219
inline void decomposition(const orc_params *p, const uint32_t *trlwe, uint32_t offset, uint32_t *dec) {
    const int L = p->L;
    const uint32_t MASK = (1u << p->bgbit) - 1u, HALF_BG = 1u << (p->bgbit - 1);
    const uint32_t *a = trlwe, *b = trlwe + N1;
    for (int j = 0; j < N1; j++) {
        const uint32_t tmp0 = a[j] + offset, tmp1 = b[j] + offset;
        for (int i = 0; i < L; i++) {
            const int sh = 32 - (i + 1) * p->bgbit;
            dec[(size_t)i * N1 + j] = ((tmp0 >> sh) & MASK) - HALF_BG;
            dec[(size_t)(i + L) * N1 + j] = ((tmp1 >> sh) & MASK) - HALF_BG;
        }
    }
}

// trgsw.zig:111-154
inline void external_product(const orc_params *p, const double *tg, const uint32_t *trlwe, uint32_t offset,
                             uint32_t *out, double *max_frac) {
    const int L2 = 2 * p->L;
    std::vector<uint32_t> dec((size_t)L2 * N1);
    decomposition(p, trlwe, offset, dec.data());
    double out_a[N1], out_b[N1];
    std::fill(out_a, out_a + N1, 0.0); std::fill(out_b, out_b + N1, 0.0);
    std::vector<double> dfft((size_t)L2 * N1);
    for (int i = 0; i < L2; i++) ifft1024(&dec[(size_t)i * N1], &dfft[(size_t)i * N1]);   // fft.zig:447-454
    for (int i = 0; i < L2; i++) {                                                          // trgsw.zig:139-142
        fma_in_fd_1024(out_a, &dfft[(size_t)i * N1], tg + ((size_t)i * 2 + 0) * N1);
        fma_in_fd_1024(out_b, &dfft[(size_t)i * N1], tg + ((size_t)i * 2 + 1) * N1);
    }
    fft1024(out_a, out, max_frac);          // trgsw.zig:145
    fft1024(out_b, out + N1, max_frac);     // trgsw.zig:146
}

// trgsw.zig:260-284
inline void cmux(const orc_params *p, const uint32_t *in1, const uint32_t *in2, const double *cond, uint32_t offset,
                 uint32_t *out, double *max_frac) {
    uint32_t tmp[2 * N1], tmp2[2 * N1];
    for (int i = 0; i < 2 * N1; i++) tmp[i] = in2[i] - in1[i];
    external_product(p, cond, tmp, offset, tmp2, max_frac);
    for (int i = 0; i < 2 * N1; i++) out[i] = tmp2[i] + in1[i];
}

// trgsw.zig:442-466
inline void poly_mul_with_xk(const uint32_t *a, size_t k, uint32_t *res) {
    const size_t N = N1;
    if (k < N) {
        std::memcpy(res + k, a, (N - k) * sizeof(uint32_t));
        for (size_t i = N - k; i < N; i++) res[i + k - N] = 0u - a[i];
    } else {
        for (size_t i = 0; i < 2 * N - k; i++) res[i + k - N] = 0u - a[i];
        std::memcpy(res, a + (2 * N - k), (N - (2 * N - k)) * sizeof(uint32_t));
    }
}

inline void gen_testvec(uint32_t *tv) {     // key.zig:134-145
    const uint32_t b_torus = f64_to_torus(0.125);
    for (int i = 0; i < N1; i++) { tv[i] = 0; tv[N1 + i] = b_torus; }
}

// trgsw.zig:290-333 (testvec == cloud key's) and 336-400 (custom)
inline void blind_rotate(const orc_params *p, const uint32_t *src, const double *bsk, uint32_t offset,
                         const uint32_t *testvec, uint32_t *out, uint32_t *trace, double *max_frac) {
    uint32_t tvbuf[2 * N1];
    if (!testvec) { gen_testvec(tvbuf); testvec = tvbuf; }
    const int n = p->n;
    const size_t row = (size_t)2 * p->L * 2 * N1;
    const size_t b_tilda = 2 * (size_t)N1 -
        (((size_t)src[n] + ((size_t)1 << (TORUS_SIZE - 1 - NBIT - 1))) >> (TORUS_SIZE - NBIT - 1));   // trgsw.zig:297
    uint32_t result[2 * N1], res2[2 * N1], nxt[2 * N1];
    poly_mul_with_xk(testvec, b_tilda, result);                 // trgsw.zig:300
    poly_mul_with_xk(testvec + N1, b_tilda, result + N1);       // trgsw.zig:301
    for (int i = 0; i < n; i++) {
        const size_t a_tilda = ((size_t)src[i] + ((size_t)1 << (TORUS_SIZE - 1 - NBIT - 1))) >> (TORUS_SIZE - NBIT - 1);  // :312
        poly_mul_with_xk(result, a_tilda, res2);                // trgsw.zig:315
        poly_mul_with_xk(result + N1, a_tilda, res2 + N1);      // trgsw.zig:316
        cmux(p, result, res2, bsk + (size_t)i * row, offset, nxt, max_frac);   // trgsw.zig:323
        std::memcpy(result, nxt, sizeof(result));
        if (trace) std::memcpy(trace + (size_t)i * 2 * N1, result, sizeof(result));
    }
    std::memcpy(out, result, sizeof(result));
}

// trlwe.zig:146-162
inline void sample_extract_index(const uint32_t *trlwe, int k, uint32_t *out, int len) {
    for (int i = 0; i < len; i++) {
        if (i <= k) out[i] = trlwe[k - i];
        else out[i] = 0u - trlwe[N1 + k - i];
    }
    out[len] = trlwe[N1 + k];
}

// trgsw.zig:471-502
inline void identity_key_switching(const orc_params *p, const uint32_t *src, const uint32_t *ksk, uint32_t *res) {
    const int n = p->n, BASEBIT = p->basebit, IKS_T = p->iks_t;
    const size_t BASE = (size_t)1 << BASEBIT;
    const size_t rows = BASE * IKS_T * N1;
    for (int x = 0; x < n; x++) res[x] = 0;
    res[n] = src[N1];                                                     // trgsw.zig:481
    const uint32_t PREC_OFFSET = 1u << (32 - (1 + BASEBIT * IKS_T));      // trgsw.zig:483
    for (int i = 0; i < N1; i++) {
        const uint32_t a_bar = src[i] + PREC_OFFSET;
        for (int j = 0; j < IKS_T; j++) {
            const int sh = 32 - (j + 1) * BASEBIT;
            const uint32_t k = (a_bar >> sh) & (uint32_t)(BASE - 1);
            if (k != 0) {
                const size_t idx = (BASE * IKS_T * (size_t)i) + (BASE * (size_t)j) + k;
                if (idx < rows) {
                    const uint32_t *row = ksk + idx * (size_t)(n + 1);
                    for (int x = 0; x <= n; x++) res[x] = res[x] - row[x];
                }
            }
        }
    }
}

// gates.zig:48-121
inline void gate_linear(const orc_params *p, int op, const uint32_t *a, const uint32_t *b, uint32_t *lin) {
    const int len = p->n + 1;
    double c = 0.0;
    for (int i = 0; i < len; i++) {
        switch (op) {
            case ORC_NAND:  lin[i] = (0u - a[i]) + (0u - b[i]); c = 0.125; break;    // :49-52
            case ORC_OR:    lin[i] = a[i] + b[i]; c = 0.125; break;                  // :58-59
            case ORC_AND:   lin[i] = a[i] + b[i]; c = -0.125; break;                 // :65-66
            case ORC_XOR:   lin[i] = a[i] + b[i] * 2u; c = 0.25; break;              // :72-73 (addMul)
            case ORC_XNOR:  lin[i] = a[i] - b[i] * 2u; c = -0.25; break;             // :79-80 (subMul)
            case ORC_NOR:   lin[i] = (0u - a[i]) + (0u - b[i]); c = -0.125; break;   // :86-89
            case ORC_ANDNY: lin[i] = (0u - a[i]) + b[i]; c = -0.125; break;          // :95-97
            case ORC_ANDYN: lin[i] = a[i] - b[i]; c = -0.125; break;                 // :103-104
            case ORC_ORNY:  lin[i] = (0u - a[i]) + b[i]; c = 0.125; break;           // :110-112
            case ORC_ORYN:  lin[i] = a[i] - b[i]; c = 0.125; break;                  // :118-119
            default: lin[i] = 0; break;
        }
    }
    lin[len - 1] += f64_to_torus(c);
}

inline void bootstrap(const orc_params *p, const uint32_t *ct, const double *bsk, const uint32_t *ksk,
                      uint32_t offset, const uint32_t *testvec, uint32_t *out) {
    uint32_t acc[2 * N1], lv1[N1 + 1];
    blind_rotate(p, ct, bsk, offset, testvec, acc, nullptr, nullptr);   // vanilla.zig:43
    sample_extract_index(acc, 0, lv1, N1);                              // vanilla.zig:47
    identity_key_switching(p, lv1, ksk, out);                           // vanilla.zig:51
}

// tlwe.zig:34-50
inline void tlwe_encrypt_f64(Rng &rng, double mu, double alpha, const uint32_t *key, int keylen, uint32_t *out) {
    uint32_t inner = 0;
    for (int i = 0; i < keylen; i++) {
        const uint32_t r = rng.u32();
        inner += key[i] * r;
        out[i] = r;
    }
    NormalDist nd(0.0, alpha);
    out[keylen] = inner + gaussian_f64(mu, nd, rng);
}

// trlwe.zig:30-64
inline void trlwe_encrypt_f64(Rng &rng, const double *mu, double alpha, const uint32_t *s1, uint32_t *out) {
    uint32_t *a = out, *b = out + N1;
    for (int i = 0; i < N1; i++) a[i] = rng.u32();
    NormalDist nd(0.0, alpha);
    for (int i = 0; i < N1; i++) b[i] = gaussian_torus(f64_to_torus(mu ? mu[i] : 0.0), nd, rng);   // utils.zig:107-118
    uint32_t pr[N1];
    poly_mul_fft(a, s1, pr);
    for (int i = 0; i < N1; i++) b[i] += pr[i];
}

// trgsw.zig:35-71 then 81-91
inline void trgsw_encrypt_fft(const orc_params *p, Rng &rng, uint32_t msg, double alpha, const uint32_t *s1, double *out,
                              bool noiseless) {
    const int L = p->L;
    std::vector<uint32_t> t((size_t)2 * L * 2 * N1);
    for (int i = 0; i < 2 * L; i++) {
        if (!noiseless) trlwe_encrypt_f64(rng, nullptr, alpha, s1, &t[(size_t)i * 2 * N1]);
        else {  // zero-noise TRLWE(0): b = a*s exactly (integer schoolbook, no FFT rounding)
            uint32_t *a = &t[(size_t)i * 2 * N1], *b = a + N1;
            for (int k = 0; k < N1; k++) a[k] = rng.u32();
            orc_poly_mul_naive(a, s1, b, N1);
        }
    }
    for (int i = 0; i < L; i++) {
        const double pf = std::pow((double)(1u << p->bgbit), -(double)(i + 1));   // trgsw.zig:48
        const uint32_t pt = f64_to_torus(pf);
        t[(size_t)i * 2 * N1 + 0] += msg * pt;                  // trlwe[i].a[0], trgsw.zig:66
        t[(size_t)(i + L) * 2 * N1 + N1 + 0] += msg * pt;       // trlwe[i+L].b[0], trgsw.zig:67
    }
    for (int i = 0; i < 2 * L; i++) {                           // trlwe.zig:111-132
        ifft1024(&t[(size_t)i * 2 * N1], out + ((size_t)i * 2 + 0) * N1);
        ifft1024(&t[(size_t)i * 2 * N1 + N1], out + ((size_t)i * 2 + 1) * N1);
    }
}

template <class F>
void par_for(size_t B, int nthreads, F f) {
    if (nthreads <= 1 || B <= 1) { for (size_t i = 0; i < B; i++) f(i); return; }
    nthreads = (int)std::min<size_t>(nthreads, B);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        const size_t lo = B * t / nthreads, hi = B * (t + 1) / nthreads;   // static contiguous partition
        th.emplace_back([=]() { for (size_t i = lo; i < hi; i++) f(i); });
    }
    for (auto &x : th) x.join();
}

}  // namespace

extern "C" {

int orc_get_params(const char *name, orc_params *o) {
    struct E { const char *nm; orc_params p; };
    // params.zig:70-375 (n, N, L, bgbit, basebit, iks_t, alpha_lv0, alpha_lv1)
    static const E T[] = {
        {"80",    {550, 1024, 3, 6, 2, 7, 5.0e-5, 3.73e-8}},
        {"110",   {630, 1024, 3, 6, 2, 8, 3.0517578125e-05, 2.9802322387695313e-8}},
        {"128",   {700, 1024, 3, 6, 2, 9, 2.0e-5, 2.0e-8}},
        {"uint1", {700, 1024, 2, 10, 2, 8, 2.0e-05, 2.0e-08}},
        {"uint2", {687, 1024, 1, 18, 4, 3, 0.00002120846893069971872305794214, 0.00000000000231841227527049948463}},
        {"uint3", {820, 1024, 1, 23, 6, 2, 0.00000251676160959795544987084234, 0.00000000000000022204460492503131}},
        {"uint4", {820, 1024, 1, 22, 5, 3, 0.00000251676160959795544987084234, 0.00000000000000022204460492503131}},
        {"uint5", {1071, 1024, 1, 22, 6, 3, 7.088226765410429399593757e-08, 2.2204460492503131e-17}},
        {"uint6", {1071, 1024, 1, 22, 6, 3, 7.088226765410429399593757e-08, 2.2204460492503131e-17}},
        {"uint7", {1160, 1024, 1, 22, 7, 3, 1.966220007498402695211596e-08, 2.2204460492503131e-17}},
        {"uint8", {1160, 1024, 1, 22, 7, 3, 1.966220007498402695211596e-08, 2.2204460492503131e-17}},
    };
    for (const E &e : T) if (!std::strcmp(e.nm, name)) { *o = e.p; return 0; }
    return -1;
}

uint32_t orc_f64_to_torus(double d) { return f64_to_torus(d); }
double orc_torus_to_f64(uint32_t t) { return (double)t / 4294967296.0; }

void orc_ifft1024(const uint32_t *in, double *out) { ifft1024(in, out); }
void orc_fft1024(const double *in, uint32_t *out) { fft1024(in, out, nullptr); }
void orc_fft1024_margin(const double *in, uint32_t *out, double *max_frac) { fft1024(in, out, max_frac); }
void orc_poly_mul_fft(const uint32_t *a, const uint32_t *b, uint32_t *out) { poly_mul_fft(a, b, out); }

void orc_poly_mul_naive(const uint32_t *a, const uint32_t *b, uint32_t *res, int n) {
    std::vector<uint32_t> r((size_t)n, 0u);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i + j < n) r[i + j] += a[i] * b[j];
            else r[i + j - n] -= a[i] * b[j];
        }
    std::memcpy(res, r.data(), (size_t)n * sizeof(uint32_t));
}

// generic-N radix-2 (fft.zig:582-669) with the live recurrence, interleaved complex
void orc_radix2_fft(double *d, int n, int inverse) {
    int j = 0;
    for (int i = 0; i < n; i++) {
        if (j > i) { std::swap(d[2 * i], d[2 * j]); std::swap(d[2 * i + 1], d[2 * j + 1]); }
        int mask = n >> 1;
        while (mask > 0 && (j & mask) != 0) { j ^= mask; mask >>= 1; }
        j ^= mask;
    }
    for (int len = 2; len <= n; len *= 2) {
        const double angle = inverse ? 2.0 * M_PI / (double)len : -2.0 * M_PI / (double)len;
        const double wl_re = std::cos(angle), wl_im = std::sin(angle);
        for (int i = 0; i < n; i += len) {
            double w_re = 1.0, w_im = 0.0;
            for (int k = 0; k < len / 2; k++) {
                const double u_re = d[2 * (i + k)], u_im = d[2 * (i + k) + 1];
                const double x_re = d[2 * (i + k + len / 2)], x_im = d[2 * (i + k + len / 2) + 1];
                const double v_re = x_re * w_re - x_im * w_im, v_im = x_re * w_im + x_im * w_re;
                d[2 * (i + k)] = u_re + v_re; d[2 * (i + k) + 1] = u_im + v_im;
                d[2 * (i + k + len / 2)] = u_re - v_re; d[2 * (i + k + len / 2) + 1] = u_im - v_im;
                const double temp = w_re * wl_re - w_im * wl_im;
                w_im = w_re * wl_im + w_im * wl_re;
                w_re = temp;
            }
        }
    }
}

size_t orc_bsk_len(const orc_params *p) { return (size_t)p->n * 2 * p->L * 2 * N1; }
size_t orc_ksk_len(const orc_params *p) { return (size_t)N1 * p->iks_t * ((size_t)1 << p->basebit) * (p->n + 1); }

void orc_gen_secret(const orc_params *p, uint64_t seed, uint32_t *s0, uint32_t *s1) {
    Rng rng(seed);
    for (int i = 0; i < p->n; i++) s0[i] = rng.boolean() ? 1u : 0u;   // key.zig:49-51
    for (int i = 0; i < N1; i++) s1[i] = rng.boolean() ? 1u : 0u;     // key.zig:52-54
}

uint32_t orc_decomposition_offset(const orc_params *p) {              // key.zig:121-131
    uint32_t offset = 0;
    const uint32_t BG = 1u << p->bgbit;
    for (int i = 0; i < p->L; i++) {
        const int sh = TORUS_SIZE - (i + 1) * p->bgbit;
        offset += (BG / 2) * (1u << sh);
    }
    return offset;
}

void orc_gen_testvec(const orc_params *, uint32_t *tv) { gen_testvec(tv); }

void orc_gen_ksk(const orc_params *p, uint64_t seed, const uint32_t *s0, const uint32_t *s1, uint32_t *ksk) {
    const int n = p->n, BASEBIT = p->basebit, IKS_T = p->iks_t;
    const size_t BASE = (size_t)1 << BASEBIT;
    std::memset(ksk, 0, orc_ksk_len(p) * sizeof(uint32_t));
    const int nt = std::max(1, orc_hardware_threads());
    par_for((size_t)N1, nt, [&](size_t i) {
        Rng rng(seed * 0x9e3779b97f4a7c15ULL + 0x1000000ULL + i);
        for (int j = 0; j < IKS_T; j++)
            for (size_t k = 1; k < BASE; k++) {                       // key.zig:158-169 (k == 0 skipped)
                const int shift_amount = (j + 1) * BASEBIT;
                const double pv = ((double)k * (double)s1[i]) / (double)((uint32_t)1 << shift_amount);   // key.zig:164
                const size_t idx = (BASE * IKS_T * i) + (BASE * j) + k;
                tlwe_encrypt_f64(rng, pv, p->alpha_lv0, s0, n, ksk + idx * (size_t)(n + 1));
            }
    });
}

static void gen_bsk_impl(const orc_params *p, uint64_t seed, const uint32_t *s0, const uint32_t *s1, double *bsk, bool noiseless) {
    const size_t row = (size_t)2 * p->L * 2 * N1;
    plan();
    const int nt = std::max(1, orc_hardware_threads());
    par_for((size_t)p->n, nt, [&](size_t i) {
        Rng rng(seed * 0x9e3779b97f4a7c15ULL + 0x2000000ULL + i);
        trgsw_encrypt_fft(p, rng, s0[i], p->alpha_lv1, s1, bsk + i * row, noiseless);   // key.zig:197-208
    });
}
void orc_gen_bsk(const orc_params *p, uint64_t seed, const uint32_t *s0, const uint32_t *s1, double *bsk) {
    gen_bsk_impl(p, seed, s0, s1, bsk, false);
}
void orc_gen_bsk_noiseless(const orc_params *p, uint64_t seed, const uint32_t *s0, const uint32_t *s1, double *bsk) {
    gen_bsk_impl(p, seed, s0, s1, bsk, true);
}

void orc_tlwe_encrypt_f64(const orc_params *, uint64_t seed, double mu, double alpha, const uint32_t *key, int keylen, uint32_t *out) {
    Rng rng(seed);
    tlwe_encrypt_f64(rng, mu, alpha, key, keylen, out);
}

void orc_tlwe_encrypt_bools(const orc_params *p, uint64_t seed, const uint8_t *bits, size_t B, const uint32_t *s0, uint32_t *out) {
    const int n = p->n;
    par_for(B, std::max(1, orc_hardware_threads()), [&](size_t i) {
        Rng rng(seed * 0x9e3779b97f4a7c15ULL + 0x3000000ULL + i);
        tlwe_encrypt_f64(rng, bits[i] ? 0.125 : -0.125, p->alpha_lv0, s0, n, out + i * (size_t)(n + 1));   // tlwe.zig:53-56
    });
}

uint32_t orc_tlwe_phase(const uint32_t *ct, const uint32_t *key, int keylen) {
    uint32_t inner = 0;
    for (int i = 0; i < keylen; i++) inner += ct[i] * key[i];
    return ct[keylen] - inner;
}

void orc_tlwe_decrypt_bools(const orc_params *, const uint32_t *ct, size_t B, const uint32_t *key, int keylen, uint8_t *bits) {
    for (size_t i = 0; i < B; i++)
        bits[i] = ((int32_t)orc_tlwe_phase(ct + i * (size_t)(keylen + 1), key, keylen) >= 0) ? 1 : 0;   // tlwe.zig:59-69
}

void orc_tlwe_encrypt_lwe_messages(const orc_params *p, uint64_t seed, const uint32_t *msgs, size_t B, uint32_t modulus,
                                   const uint32_t *s0, uint32_t *out) {
    const int n = p->n;
    par_for(B, std::max(1, orc_hardware_threads()), [&](size_t i) {
        Rng rng(seed * 0x9e3779b97f4a7c15ULL + 0x4000000ULL + i);
        const uint32_t nm = msgs[i] % modulus;                              // tlwe.zig:81
        const double scale = 1.0 / (2.0 * (double)modulus);                 // tlwe.zig:84
        tlwe_encrypt_f64(rng, (double)nm * scale, p->alpha_lv0, s0, n, out + i * (size_t)(n + 1));
    });
}

void orc_tlwe_decrypt_lwe_messages(const orc_params *, const uint32_t *ct, size_t B, uint32_t modulus, const uint32_t *key,
                                   int keylen, uint32_t *msgs) {
    for (size_t i = 0; i < B; i++) {
        const uint32_t res = orc_tlwe_phase(ct + i * (size_t)(keylen + 1), key, keylen);
        const double res_f64 = orc_torus_to_f64(res);
        const double scale = 1.0 / (2.0 * (double)modulus);
        const uint64_t message = (uint64_t)(res_f64 / scale + 0.5);         // tlwe.zig:113
        msgs[i] = (uint32_t)(message % modulus);
    }
}

void orc_trlwe_encrypt_f64(const orc_params *, uint64_t seed, const double *mu, double alpha, const uint32_t *s1, uint32_t *out) {
    Rng rng(seed);
    trlwe_encrypt_f64(rng, mu, alpha, s1, out);
}

void orc_trlwe_phase(const orc_params *, const uint32_t *ct, const uint32_t *s1, uint32_t *phase) {
    uint32_t pr[N1];
    poly_mul_fft(ct, s1, pr);                                               // trlwe.zig:87
    for (int i = 0; i < N1; i++) phase[i] = ct[N1 + i] - pr[i];             // trlwe.zig:93
}

void orc_trgsw_encrypt_fft(const orc_params *p, uint64_t seed, uint32_t msg, double alpha, const uint32_t *s1, double *out) {
    Rng rng(seed);
    trgsw_encrypt_fft(p, rng, msg, alpha, s1, out, false);
}

void orc_gate_linear(const orc_params *p, int op, const uint32_t *a, const uint32_t *b, uint32_t *lin) { gate_linear(p, op, a, b, lin); }
void orc_gate_not(const orc_params *p, const uint32_t *a, uint32_t *out) { for (int i = 0; i <= p->n; i++) out[i] = 0u - a[i]; }
void orc_gate_constant(const orc_params *p, int value, uint32_t *out) {
    uint32_t mu = f64_to_torus(0.125);
    mu = value ? mu : (1u - mu);                                            // gates.zig:146-147 (quirk kept)
    for (int i = 0; i < p->n; i++) out[i] = 0;
    out[p->n] = mu;
}
void orc_poly_mul_with_xk(const uint32_t *a, size_t k, uint32_t *res) { poly_mul_with_xk(a, k, res); }
void orc_decomposition(const orc_params *p, const uint32_t *trlwe, uint32_t offset, uint32_t *dec) { decomposition(p, trlwe, offset, dec); }
void orc_external_product(const orc_params *p, const double *tg, const uint32_t *trlwe, uint32_t offset, uint32_t *out, double *max_frac) {
    external_product(p, tg, trlwe, offset, out, max_frac);
}

void orc_external_product_int(const orc_params *p, const double *tg, const uint32_t *trlwe, uint32_t offset, uint32_t *out) {
    const int L2 = 2 * p->L;
    std::vector<uint32_t> dec((size_t)L2 * N1);
    decomposition(p, trlwe, offset, dec.data());
    std::vector<uint32_t> acc(2 * N1, 0u), Bt(N1), pr(N1);
    for (int i = 0; i < L2; i++)
        for (int ab = 0; ab < 2; ab++) {
            fft1024(tg + ((size_t)i * 2 + ab) * N1, Bt.data(), nullptr);   // recover time-domain row (exact round trip)
            // negacyclic schoolbook with signed digits, mod 2^32
            std::vector<int64_t> r(N1, 0);
            for (int x = 0; x < N1; x++) {
                const int64_t d = (int32_t)dec[(size_t)i * N1 + x];
                if (d == 0) continue;
                for (int y = 0; y < N1; y++) {
                    const int64_t v = d * (int64_t)(int32_t)Bt[y];
                    if (x + y < N1) r[x + y] += v; else r[x + y - N1] -= v;
                }
            }
            for (int x = 0; x < N1; x++) acc[(size_t)ab * N1 + x] += (uint32_t)(uint64_t)r[x];
        }
    std::memcpy(out, acc.data(), 2 * N1 * sizeof(uint32_t));
}

void orc_cmux(const orc_params *p, const uint32_t *in1, const uint32_t *in2, const double *cond, uint32_t offset, uint32_t *out, double *max_frac) {
    cmux(p, in1, in2, cond, offset, out, max_frac);
}
void orc_blind_rotate(const orc_params *p, const uint32_t *src, const double *bsk, uint32_t offset, const uint32_t *testvec,
                      uint32_t *out, uint32_t *trace, double *max_frac) {
    blind_rotate(p, src, bsk, offset, testvec, out, trace, max_frac);
}
void orc_sample_extract_index(const orc_params *, const uint32_t *trlwe, int k, uint32_t *out) { sample_extract_index(trlwe, k, out, N1); }
void orc_sample_extract_index2(const orc_params *p, const uint32_t *trlwe, int k, uint32_t *out) { sample_extract_index(trlwe, k, out, p->n); }
void orc_identity_key_switching(const orc_params *p, const uint32_t *lv1, const uint32_t *ksk, uint32_t *out) { identity_key_switching(p, lv1, ksk, out); }
void orc_bootstrap(const orc_params *p, const uint32_t *ct, const double *bsk, const uint32_t *ksk, uint32_t offset,
                   const uint32_t *testvec, uint32_t *out) {
    bootstrap(p, ct, bsk, ksk, offset, testvec, out);
}

void orc_gate_batch(const orc_params *p, int op, const int32_t *ops, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t B,
                    const double *bsk, const uint32_t *ksk, uint32_t offset, int nthreads) {
    const size_t w = (size_t)p->n + 1;
    plan();
    par_for(B, nthreads, [&](size_t i) {
        std::vector<uint32_t> lin(w);
        gate_linear(p, ops ? ops[i] : op, a + i * w, b + i * w, lin.data());
        bootstrap(p, lin.data(), bsk, ksk, offset, nullptr, out + i * w);
    });
}

void orc_bootstrap_batch(const orc_params *p, const uint32_t *in, uint32_t *out, size_t B, const double *bsk, const uint32_t *ksk,
                         uint32_t offset, const uint32_t *testvec, int tv_per_item, int nthreads) {
    const size_t w = (size_t)p->n + 1;
    plan();
    par_for(B, nthreads, [&](size_t i) {
        const uint32_t *tv = testvec ? testvec + (tv_per_item ? i * 2 * N1 : 0) : nullptr;
        bootstrap(p, in + i * w, bsk, ksk, offset, tv, out + i * w);
    });
}

void orc_blind_rotate_batch(const orc_params *p, const uint32_t *in, uint32_t *out, size_t B, const double *bsk, uint32_t offset,
                            const uint32_t *testvec, int tv_per_item, int nthreads) {
    const size_t w = (size_t)p->n + 1;
    plan();
    par_for(B, nthreads, [&](size_t i) {
        const uint32_t *tv = testvec ? testvec + (tv_per_item ? i * 2 * N1 : 0) : nullptr;
        blind_rotate(p, in + i * w, bsk, offset, tv, out + i * 2 * N1, nullptr, nullptr);
    });
}

void orc_keyswitch_batch(const orc_params *p, const uint32_t *lv1, uint32_t *lv0, size_t B, const uint32_t *ksk, int nthreads) {
    par_for(B, nthreads, [&](size_t i) {
        identity_key_switching(p, lv1 + i * (size_t)(N1 + 1), ksk, lv0 + i * (size_t)(p->n + 1));
    });
}

void orc_gen_reenc_key(const orc_params *p, uint64_t seed, const uint32_t *key_from, const uint32_t *key_to, int basebit, int t, uint32_t *out) {
    const int n = p->n;
    const size_t base = (size_t)1 << basebit;
    std::memset(out, 0, (size_t)n * t * base * (n + 1) * sizeof(uint32_t));      // proxy_reenc.zig:220-222
    par_for((size_t)n, std::max(1, orc_hardware_threads()), [&](size_t i) {
        Rng rng(seed * 0x9e3779b97f4a7c15ULL + 0x5000000ULL + i);
        for (int j = 0; j < t; j++)
            for (size_t k = 1; k < base; k++) {                                   // proxy_reenc.zig:225-241
                const double pv = ((double)k * (double)key_from[i]) / (double)((uint32_t)1 << ((j + 1) * basebit));
                const size_t idx = (base * t * i) + (base * j) + k;
                tlwe_encrypt_f64(rng, pv, p->alpha_lv0, key_to, n, out + idx * (size_t)(n + 1));
            }
    });
}

void orc_reencrypt(const orc_params *p, const uint32_t *ct, const uint32_t *key, int basebit, int t, uint32_t *res) {
    const int n = p->n;
    const size_t base = (size_t)1 << basebit;
    for (int x = 0; x < n; x++) res[x] = 0;
    res[n] = ct[n];                                                               // proxy_reenc.zig:276
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));                  // proxy_reenc.zig:279
    for (int i = 0; i < n; i++) {
        const uint32_t a_bar = ct[i] + prec_offset;
        for (int j = 0; j < t; j++) {
            const uint32_t k = (a_bar >> (32 - (j + 1) * basebit)) & (uint32_t)(base - 1);
            if (k != 0) {
                const uint32_t *row = key + ((base * t * (size_t)i) + (base * (size_t)j) + k) * (size_t)(n + 1);
                for (int x = 0; x <= n; x++) res[x] -= row[x];                    // proxy_reenc.zig:299-301
            }
        }
    }
}

uint32_t orc_lut_encode(uint32_t message, uint32_t modulus) {               // encoder.zig:66-73
    const double scale = 1.0 / (2.0 * (double)modulus);                     // encoder.zig:35
    return f64_to_torus((double)(message % modulus) * scale);
}
uint32_t orc_lut_decode(uint32_t value, uint32_t modulus) {                 // encoder.zig:96-105
    const double scale = 1.0 / (2.0 * (double)modulus);
    const uint64_t m = (uint64_t)(orc_torus_to_f64(value) / scale + 0.5);
    return (uint32_t)(m % modulus);
}

static size_t div_round(size_t a, size_t b) { return (a + b / 2) / b; }     // generator.zig:253-255

void orc_lut_generate(const orc_params *, const uint32_t *table, uint32_t modulus, uint32_t *out) {
    const size_t sz = N1;
    std::vector<uint32_t> raw(sz, 0u), rot(sz, 0u);
    for (size_t x = 0; x < modulus; x++) {                                  // generator.zig:95-110
        const size_t start = div_round(x * sz, modulus), end = div_round((x + 1) * sz, modulus);
        const uint32_t enc = orc_lut_encode(table[x], modulus);
        for (size_t xx = start; xx < end; xx++) raw[xx] = enc;
    }
    const size_t offset = div_round(sz, 2 * (size_t)modulus);               // generator.zig:113
    for (size_t i = 0; i < sz; i++) rot[i] = raw[(i + offset) % sz];        // generator.zig:120-123
    for (size_t i = sz - offset; i < sz; i++) rot[i] = ~rot[i] + 1u;        // generator.zig:126-128
    for (size_t i = 0; i < sz; i++) { out[i] = 0; out[N1 + i] = rot[i]; }   // generator.zig:131-134
}

int orc_hardware_threads(void) {
    const char *e = std::getenv("ORC_THREADS");
    if (e && *e) return std::max(1, std::atoi(e));
    unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

}  // extern "C"

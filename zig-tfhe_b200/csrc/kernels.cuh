// kernels.cuh -- launch interfaces of the device kernels (internal to libtfhe_b200).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "negacyclic_fft.cuh"

namespace tfhe_b200 {

// Arguments of the blind-rotation kernel K1 (one 64-thread group per ciphertext).
struct BrArgs {
    const uint32_t *in_a;   // [B][n+1]: first gate operand, or the ciphertext itself when op < 0
    const uint32_t *in_b;   // [B][n+1]: second gate operand (ignored when the opcode is < 0)
    const int32_t *ops;     // per-item opcode, or nullptr -> `op`
    // circuit level (capi.cu, tfhe_b200_circuit_run): when lvl_a != nullptr, item ct = gate slot g * inst + instance k reads
    // its operands from wire rows of in_a: row = (lvl_a[g] & 0x7fffffff) * inst + k; bit 31 = Gates.notGate of that wire
    const int32_t *lvl_ops;
    const uint32_t *lvl_a, *lvl_b;
    uint32_t inst;
    int op;                 // tfhe_b200_gate, or -1 = plain bootstrap input (no linear part)
    const cplx *bsk;        // device layout: [n*2L] chunks of [ab][q0][t] cplx (16 KiB each)
    const cplx *tw2;        // [7][8]
    const cplx *tw3;        // [7][64]
    const uint32_t *testvec;  // nullptr -> (a = 0, b = 2^29); else [2][N] or [B][2][N]
    int tv_per_item;
    uint32_t *out_lv1;      // [B][N+1] sampleExtractIndex(.,0), or nullptr
    uint32_t *out_trlwe;    // [B][2][N] accumulator, or nullptr
    unsigned long long *margin_bits;  // global max |t-round(t)| as double bits (MARGIN variant), or nullptr
    uint32_t B;             // ciphertexts of THIS launch
    uint32_t ct_base;       // index of its first ciphertext in the arrays above (a batch split into several launches)
    int n, L, bgbit;
    uint32_t offset;        // CloudKey.decomposition_offset
    int wide_round;         // 1: F2I.S64 rounding (large-digit sets), 0: magic-add rounding
    int ms_shift;           // modulus switch to multiples of 2^ms_shift (0: the reference's; > 0: many-function bootstrap)
#ifdef TFHE_B200_DIAG
    int diag;               // diagnostic builds only (blind_rotate.cu): parts of the kernel switched off
#endif
};

struct BrTuning {
    int kct = 0;        // ciphertexts per CTA (0 = default)
    int use_tma = 1;    // stream key chunks with cp.async.bulk + mbarrier (0: direct global loads)
    int sm_count = 0;   // SMs of the target device (wave-quantisation aware choice of kct)
    int latency_mode = 1;   // batches <= sm_count: one CTA per ciphertext, transforms of an iteration in parallel; batches <=
                            // sm_count / 2: one two-CTA cluster per ciphertext (2: never the cluster kernel, 0: throughput kernel only)
    int concurrent = 0; // 1: other kernels share the GPU (circuit lanes): pick the CTA width by work per SM-second, not by waves
    int team = 0;       // ciphertexts sharing a warp in adjacent lanes (0 = default, 1, 2)
    int twt = 0;        // pass-2 / pass-3 twiddles in tensor memory: 0 = where it wins (six ciphertexts per CTA), 1 = also at 4 and 5, -1 = never
    int diag = 0;       // diagnostic builds only
};

// returns cudaSuccess or the launch error; *launches += kernels launched
cudaError_t launch_blind_rotate(const BrArgs &a, const BrTuning &tune, bool track_margin, cudaStream_t s, uint64_t *launches);
// Exact mode.  tables: make_exact_tables() image (6 x 512 doubles); bsk_ref: key in the reference layout (legacy kernel);
// bsk_x: key in the exact chunk layout (launch_permute_bsk_exact); shared_tab: make_exact_shared_tables() image
// (512 + 7 cplx).  legacy != 0: the one-CTA-per-ciphertext shared-memory kernel of round 1 (kept for A/B and as the fallback
// when the host's libm does not give conjugate-symmetric twiddle tables).
struct ExactArgs {
    const double *tables;
    const double *bsk_ref;
    const cplx *bsk_x;
    const cplx *shared_tab;
    int legacy;
    int kct;        // ciphertexts per CTA (0 = automatic)
    int sm_count;
};
cudaError_t launch_blind_rotate_exact(const BrArgs &a, const ExactArgs &x, bool track_margin, cudaStream_t s, uint64_t *launches);
cudaError_t launch_permute_bsk_exact(const double *ref_bsk, cplx *out, int n, int L, cudaStream_t s, uint64_t *launches);

// K2: identity key switching, lv1 [B][N+1] -> lv0 [B][n+1].  ksk_dev: [N][t][base-1][pitch] u32.
struct KsArgs {
    const uint32_t *lv1;
    uint32_t *lv0;
    const uint32_t *ksk;
    uint32_t B;
    int n, basebit, iks_t, pitch;   // pitch = row length in u32 (multiple of 4)
    int in_dim;                     // mask length of the source samples: N (key switch) or n (proxy re-encryption)
    int tile = 0, vec = 0;          // tuning overrides (0 = automatic): ciphertexts per CTA, uint4 vectors per thread
    int fill = 0;                   // tuning override: CTAs per SM the i-range split aims for (0 = automatic)
    int rot = 0;                    // -1: every CTA walks i from the start of its range (default: staggered starting points)
};
cudaError_t launch_keyswitch(const KsArgs &a, int sm_count, cudaStream_t s, uint64_t *launches);

// K2t (keyswitch_tc.cu): the same key switch as an unsigned 8-bit tensor-core contraction (tcgen05.mma kind::i8) over the
// one-hot expansion of the digits; BASEBIT = 2, 4, 5 sets, in_dim = N.  ksk_tc: launch_ksk_to_tc() image of the packed key
// (keyswitch_tc_key_bytes()); digits: scratch of B * keyswitch_tc_digit_words() u64.
bool keyswitch_tc_supported(int basebit, int iks_t, int in_dim, int pitch);
size_t keyswitch_tc_key_bytes(int pitch, int basebit, int iks_t);
size_t keyswitch_tc_digit_words(int basebit, int iks_t);      // 64-bit digit words (= K blocks) per ciphertext; 0 if unsupported
cudaError_t launch_ksk_to_tc(const uint32_t *ksk_packed, uint8_t *out, int basebit, int iks_t, int pitch, cudaStream_t s, uint64_t *launches);
cudaError_t launch_keyswitch_tc(const KsArgs &a, const uint8_t *ksk_tc, uint64_t *digits, cudaStream_t s, uint64_t *launches);

// one-time key re-layout kernels
cudaError_t launch_permute_bsk(const double *ref_bsk, cplx *out, int n, int L, cudaStream_t s, uint64_t *launches);
cudaError_t launch_repack_ksk(const uint32_t *ref_ksk, size_t ref_row_stride_u32, uint32_t *out, int n, int basebit, int iks_t,
                              int pitch, int in_dim, cudaStream_t s, uint64_t *launches);
// test vectors [count][2][N] from function tables [count][m] of torus values (lut/generator.zig:158-191)
cudaError_t launch_build_testvec(const uint32_t *tables, int m, uint32_t *out, size_t count, cudaStream_t s, uint64_t *launches);
// cloud-key generation on the device (keygen.cu): key.genKeySwitchingKey / genBootstrappingKey (src/key.zig:148-212)
cudaError_t launch_keygen_ksk(const uint32_t *s0, const uint32_t *s1, uint64_t seed, double alpha, int n, int basebit, int iks_t, int pitch,
                              uint32_t *dev, uint32_t *ref, cudaStream_t s, uint64_t *launches);
cudaError_t launch_keygen_bsk(const uint32_t *s0, const uint32_t *s1, uint64_t seed, double alpha, int n, int L, int bgbit, const cplx *tw2,
                              const cplx *tw3, cplx *dev, double *ref, cudaStream_t s, uint64_t *launches);
// K3: out = -a over [B][n+1]
cudaError_t launch_negate(const uint32_t *a, uint32_t *out, size_t count, cudaStream_t s, uint64_t *launches);
// n_rows trivial ciphertexts (mask 0, body `body`) of w words each: Gates.constant (src/gates.zig:144-151)
cudaError_t launch_fill_constant(uint32_t *rows, size_t n_rows, int w, uint32_t body, cudaStream_t s, uint64_t *launches);
// trlwe.sampleExtractIndex(., k) (src/trlwe.zig:146-162) over a batch: [B][2][N] -> [B][N+1]
cudaError_t launch_sample_extract(const uint32_t *trlwe, uint32_t *lv1, uint32_t B, int k, cudaStream_t s, uint64_t *launches);
// first n entries + body of an lv1 sample -> TLWELv0-shaped "hybrid" sample (trlwe.zig:165-180)
cudaError_t launch_extract2(const uint32_t *lv1, uint32_t *out, uint32_t B, int n, cudaStream_t s, uint64_t *launches);

// FP64 FMA throughput microbenchmark (roofline denominator): returns seconds for `flops_out` flops
cudaError_t run_fp64_peak(int sm_count, cudaStream_t s, double *tflops, uint64_t *launches);

}  // namespace tfhe_b200

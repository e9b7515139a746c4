// blind_rotate.cu -- K1: fused gate-linear-part + blind rotation + sample extraction for sm_100a.
//
// Replaces, for a batch of ciphertexts at once:
//   Gates.*Gate linear parts            src/gates.zig:48-121   (fused prologue)
//   trgsw.blindRotate / WithTestvec     src/trgsw.zig:290-400
//     polyMulWithXK                     src/trgsw.zig:442-466  (index arithmetic, never materialised)
//     cmux                              src/trgsw.zig:260-284
//     externalProductWithFft            src/trgsw.zig:111-154
//       decompositionIntoStorage        src/trgsw.zig:193-219  (digits produced in registers)
//       ifft1024 x 2L, fmaInFd1024, fft1024 x 2   src/fft.zig:293-443, src/trgsw.zig:157-189
//   trlwe.sampleExtractIndex(., 0)      src/trlwe.zig:146-162  (fused epilogue)
//
// Execution model.  One CTA = KCT groups of 64 threads.  Default KCT = 4: 8 warps, 2 per SM sub-partition, so the
// per-thread twiddles of passes 2 and 3 live in registers (252-register budget) instead of being re-read from
// shared memory (the shared-memory data pipe is this kernel's limiter, profiles/r01_ncu_k1_v1_summary.txt).
// KCT = 5, 6 (168 registers) keep only r, r^2, r^4 of pass 2 and expand the other powers per pass.  A group owns one
// ciphertext for all n iterations: its TRLWE accumulator (2 x 1024 u32) never leaves shared memory,
// the 2L digit spectra never leave registers (each is consumed by the pointwise MAC as soon as its
// last radix-8 pass finishes), and the two output spectra are 16 complex accumulators per thread.
// The bootstrapping key is streamed chunk by chunk (one chunk = the a and b spectra of one gadget
// row, 16 KiB) by cp.async.bulk into an NST-deep shared-memory ring guarded by mbarriers, so the
// KCT ciphertexts of the CTA share each chunk and every CTA of the wave hits the same L2 lines.
// There is no dedicated producer warp (a 13th warp would cap the kernel at 128 registers): thread 0
// polls the ring's empty barriers at the exchange points of its own transforms and issues the next
// bulk copy as soon as a stage has been released by all groups.
//
// No tensor cores: the pointwise MAC is not a contraction and the 512-point transforms are FP64.
#include <cuda_runtime.h>

#include <algorithm>

#include "br_common.cuh"
#include "br_ring.cuh"
#include "kernels.cuh"
#include "negacyclic_fft.cuh"

namespace tfhe_b200 {

// Diagnostic builds (tools/build_diag.sh, -DTFHE_B200_DIAG): BrArgs.diag switches parts of the throughput kernel OFF to
// measure what each costs (results are then wrong on purpose).  bit 0: no key ring (no bulk copies, no mbarrier waits);
// bit 1: no key loads in the pointwise MAC; bit 2: X2 exchange without the group barrier; bit 3: no X1 exchange;
// bit 4: no X2 exchange; bit 5: the ring runs (bulk copies, releases) but nobody waits for a chunk to land.  The product
// build compiles none of this.
#ifdef TFHE_B200_DIAG
#define DIAG_ON(P, bit) (((P).diag >> (bit)) & 1)
#else
#define DIAG_ON(P, bit) 0
#endif

namespace {

// Teams of two (tuning key "team") cut the shared-memory wavefronts by 11 % (ncu) but the kernel is latency-bound,
// not shared-memory-bound: 88.8 k vs 90.8 k bootstraps/s at KCT = 4, 83.9 k at KCT = 6 with shared-memory twiddles
// (profiles/r01_team_probe.log, r01_wave_scaling.log).  The default stays one ciphertext per warp pair, KCT = 4.
// Key-ring refill policy of the throughput kernel.  true: the consumer warp that releases a stage LAST (a shared-memory
// counter per stage) issues the next bulk copy into it at once.  false (round 1, -DTFHE_B200_RING_POLL): thread 0 polls
// the stages' empty barriers at the exchange points of its own transforms -- a stage released just after a poll then
// waits most of a transform for its refill, and warp 0 runs the divergent poll code on everybody's critical path
// (profiles/r02_k1_ring.log: same speed at four ciphertexts per CTA, +6 % at six, where three warps per scheduler drift further apart).
#ifdef TFHE_B200_RING_POLL
constexpr bool kRingLastArriver = false;
#else
constexpr bool kRingLastArriver = true;
#endif
#ifdef TFHE_B200_NO_LT3
constexpr bool kUnrollL3 = false;
#else
constexpr bool kUnrollL3 = true;
#endif   // L = 3 / BGBIT = 6 instantiation of the throughput kernel (false: generic kernel only, for A/B runs)
#ifdef TFHE_B200_SEQ_INVERSE
constexpr bool kPairInverse = false;
#else
constexpr bool kPairInverse = true;
#endif   // the two inverse transforms of a step pipelined against each other (inv_transform_pair); false: one after the other

// Twiddles r^1..r^7 of one thread for one pass.  MODE 0: all seven resident (28 registers, KCT <= 4);
// MODE 1: r, r^2, r^4 resident and the rest expanded per pass (12 registers); MODE 2: read from a shared-memory
// table right before use (0 registers; with teams of two the loads of adjacent lanes merge, see the kernel).
// MODE 3: the seven twiddles live in 32 tensor-memory columns of this thread's lane (written once at kernel start) and are
// fetched with one tcgen05.ld right before each pass: 0 registers outside the pass and no shared-memory traffic, which
// is what lets 5 or 6 ciphertexts (10 / 12 warps at 200 / 168 registers) share an SM without touching the saturated
// shared-memory pipe (tuning key "twt").
constexpr int kTwFull = 0, kTwPow = 1, kTwSmem = 2, kTwTmem = 3;
template <int MODE>
struct Tw2 {
    cplx w[MODE == kTwFull ? 7 : MODE == kTwPow ? 3 : 1];
    const cplx *tab;   // MODE 2: &table[this thread's node], powers `stride` apart
    int stride;
    uint32_t taddr;    // MODE 3: tensor-memory address (lane quadrant of this warp, first of 32 columns)
    __device__ __forceinline__ void get(cplx (&out)[7]) const {
        if (MODE == kTwTmem) {
            uint32_t r[32];
            tmem_ld32(taddr, r);
            tmem_wait_ld();
#pragma unroll
            for (int p = 0; p < 7; p++) {
                out[p].re = __hiloint2double((int)r[4 * p + 1], (int)r[4 * p]);
                out[p].im = __hiloint2double((int)r[4 * p + 3], (int)r[4 * p + 2]);
            }
        } else if (MODE == kTwPow) expand_powers(out, w[0], w[1], w[2]);
        else if (MODE == kTwSmem) {
#pragma unroll
            for (int p = 0; p < 7; p++) out[p] = tab[p * stride];
        } else {
#pragma unroll
            for (int p = 0; p < 7; p++) out[p] = w[p];
        }
    }
};

// forward transform, role A registers in -> role C (leaf order) out
// (X1 WITHOUT shared memory -- the 8 x 8 transpose between the eight lanes of an octet and their registers as three butterfly
// stages of warp shuffles, 48 SHFL + 96 SEL per thread instead of 8 STS.128 + 8 LDS.128 -- was built and measured at six
// ciphertexts per CTA: bit-exact, 89.1 k instead of 103.0 k bootstraps/s.  Removed again; profiles/r02_k1_ring.log.)
// ALIAS: X1 lives in the X2 buffer this transform does NOT use for its own X2 exchange (rows of this warp only), see
// Layout::kX1Alias for why that is race-free.
template <bool USE_TMA, bool DBX2, int POW, int POW3, bool ALIAS = false>
__device__ __forceinline__ void fwd_transform(cplx (&v)[8], Xbuf &xb, const Tw2<POW> &tw2, const Tw2<POW3> &tw3, int hi, int lo,
                                              int barid, Producer &pr, int nthr = kGroupThreads, int diag = 0) {
    fwd_pass1(v);
    cplx *x1 = ALIAS ? xb.x2 + (xb.flip ^ kX2Slots) : xb.x1;
    if (!(diag & 8)) {
#pragma unroll
    for (int q = 0; q < 8; q++) x1[ALIAS ? x1a_slot(hi, q, lo) : x1_slot(hi, q, lo)] = v[q];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = x1[ALIAS ? x1a_slot(hi, lo, q) : x1_slot(hi, lo, q)];
    }
    {
        cplx w[7];
        tw2.get(w);
        fwd_pass(v, w, 1);
    }
    if (USE_TMA && !kRingLastArriver) producer_poll(pr);
    cplx *x2 = xb.x2 + (DBX2 ? xb.flip : 0);
    if (DBX2) xb.flip ^= kX2Slots;
    else bar_sync(barid, nthr);   // every reader of the previous X2 contents is done
    if (!(diag & 16)) {
#pragma unroll
    for (int q = 0; q < 8; q++) x2[x2_slot(lo, q, hi)] = v[q];
    if (diag & 4) __syncwarp(); else bar_sync(barid, nthr);
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = x2[x2_slot(hi, lo, q)];
    }
    {
        cplx w[7];
        tw3.get(w);
        fwd_pass(v, w, 1);
    }
}

// inverse transform, role C (leaf order) in -> role A out: v[p] = c_e, e = 64 p + 8 lo + hi
template <bool USE_TMA, bool DBX2, int POW, int POW3, bool ALIAS = false>
__device__ __forceinline__ void inv_transform(cplx (&v)[8], Xbuf &xb, const Tw2<POW> &tw2, const Tw2<POW3> &tw3, int hi, int lo,
                                              int barid, Producer &pr, int nthr = kGroupThreads, int diag = 0) {
    {
        cplx w[7];
        tw3.get(w);
        inv_pass(v, w, 1);
    }
    if (USE_TMA && !kRingLastArriver) producer_poll(pr);
    cplx *x2 = xb.x2 + (DBX2 ? xb.flip : 0);
    if (DBX2) xb.flip ^= kX2Slots;
    else bar_sync(barid, nthr);
    if (!(diag & 16)) {
#pragma unroll
    for (int q = 0; q < 8; q++) x2[x2_slot(hi, lo, q)] = v[q];
    if (diag & 4) __syncwarp(); else bar_sync(barid, nthr);
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = x2[x2_slot(lo, q, hi)];
    }
    {
        cplx w[7];
        tw2.get(w);
        inv_pass(v, w, 1);
    }
    // ALIAS: the inverse X2 reads cross both warps' rows, so X1 may reuse that same buffer (this warp's rows) only once
    // every thread of the group has finished them: one more group barrier, on two of the eight transforms of a step
    cplx *x1 = ALIAS ? x2 : xb.x1;
    if (ALIAS) bar_sync(barid, nthr);
    if (!(diag & 8)) {
#pragma unroll
    for (int q = 0; q < 8; q++) x1[ALIAS ? x1a_slot(hi, lo, q) : x1_slot(hi, lo, q)] = v[q];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = x1[ALIAS ? x1a_slot(hi, q, lo) : x1_slot(hi, q, lo)];
    }
    inv_pass1(v);
}

// Both inverse transforms of a CMUX step (the a and b halves of the external product) software-pipelined against each other, for the
// layout with X1 laid over the X2 buffers (Layout::kX1Alias, double-buffered X2).  Same arithmetic per polynomial as inv_transform;
// a goes through the X2 buffer `flip`, b through the other one, and each group barrier of one polynomial is covered by a pass of the
// other: three group barriers instead of four, and the shared-memory round trips of one overlap the butterflies of the other.
//   buffer F = flip: free on entry (DBX2 argument of inv_transform).  Buffer G = the last forward transform's X2: its forward reads
//   are finished by everybody once bar 1 is passed.  X1 of a reuses F after bar 2 (all X2 reads of F precede it), X1 of b reuses G
//   after bar 3.  `flip` is unchanged on exit, as after two sequential transforms.
// FIN(v, which) consumes a finished polynomial (rounding + accumulation), a's before b's last pass to free its registers.
template <int POW, int POW3, class Fin>
__device__ __forceinline__ void inv_transform_pair(cplx (&a)[8], cplx (&b)[8], Xbuf &xb, const Tw2<POW> &tw2, const Tw2<POW3> &tw3, int hi, int lo,
                                                   int barid, int nthr, Fin fin) {
    cplx *bf = xb.x2 + xb.flip, *bg = xb.x2 + (xb.flip ^ kX2Slots);
    {
        cplx w[7];
        tw3.get(w);
        inv_pass(a, w, 1);
#pragma unroll
        for (int q = 0; q < 8; q++) bf[x2_slot(hi, lo, q)] = a[q];
        inv_pass(b, w, 1);
    }
    bar_sync(barid, nthr);   // 1: a's X2 writes are visible; nobody still reads G (forward X2 of the step's last digit transform)
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = bf[x2_slot(lo, q, hi)];
#pragma unroll
    for (int q = 0; q < 8; q++) bg[x2_slot(hi, lo, q)] = b[q];
    {
        cplx w[7];
        tw2.get(w);
        inv_pass(a, w, 1);
        bar_sync(barid, nthr);   // 2: b's X2 writes are visible; everybody has finished its X2 reads of F
#pragma unroll
        for (int q = 0; q < 8; q++) b[q] = bg[x2_slot(lo, q, hi)];
#pragma unroll
        for (int q = 0; q < 8; q++) bf[x1a_slot(hi, lo, q)] = a[q];   // X1 of a in this warp's rows of F
        inv_pass(b, w, 1);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = bf[x1a_slot(hi, q, lo)];
    inv_pass1(a);
    fin(a, 0);
    bar_sync(barid, nthr);   // 3: everybody has finished its X2 reads of G
#pragma unroll
    for (int q = 0; q < 8; q++) bg[x1a_slot(hi, lo, q)] = b[q];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; q++) b[q] = bg[x1a_slot(hi, q, lo)];
    inv_pass1(b);
    fin(b, 1);
}

// 8 complex accumulators <-> 32 TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_load_cplx8(cplx (&o)[8], uint32_t taddr) {
    uint32_t r[32];
    tmem_wait_st();                 // our own earlier stores to these columns have landed
    tmem_ld32(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int q = 0; q < 8; q++) {
        o[q].re = __hiloint2double((int)r[4 * q + 1], (int)r[4 * q]);
        o[q].im = __hiloint2double((int)r[4 * q + 3], (int)r[4 * q + 2]);
    }
}
__device__ __forceinline__ void tmem_store_cplx8(uint32_t taddr, const cplx (&o)[8]) {
    uint32_t r[32];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        r[4 * q] = (uint32_t)__double2loint(o[q].re); r[4 * q + 1] = (uint32_t)__double2hiint(o[q].re);
        r[4 * q + 2] = (uint32_t)__double2loint(o[q].im); r[4 * q + 3] = (uint32_t)__double2hiint(o[q].im);
    }
    tmem_st32(taddr, r);
}

template <bool MARGIN>
__device__ __forceinline__ void round_accumulate(const cplx (&v)[8], uint32_t *accp, int t, int wide, double &margin) {
#pragma unroll
    for (int p = 0; p < 8; p++) {
        const int e = 64 * p + t;   // acc_pos of coefficient 64 p + 8 lo + hi
        if (MARGIN) {
            margin = fmax(margin, fabs(v[p].re - rint(v[p].re)));
            margin = fmax(margin, fabs(v[p].im - rint(v[p].im)));
        }
        const uint32_t r0 = wide ? round_torus_wide(v[p].re) : round_torus_magic(v[p].re);
        const uint32_t r1 = wide ? round_torus_wide(v[p].im) : round_torus_magic(v[p].im);
        // one shared-memory reduction per word instead of a load, an add and a store: half the accumulator wavefronts and no
        // load-to-store dependency (105.8 k against 104.8 k bootstraps/s at six per CTA; -DTFHE_B200_RMW_ACC builds the old form)
#ifdef TFHE_B200_RMW_ACC
        accp[e] += r0;
        accp[e + kHalfN] += r1;
#else
        atomicAdd(&accp[e], r0);
        atomicAdd(&accp[e + kHalfN], r1);
#endif
    }
}

// shared-memory footprint of one ciphertext group
template <int KCT, int TEAM = 1, bool ALIAS = false>
struct Layout {
    // ALIAS (six ciphertexts per CTA with tensor-memory twiddles): X1 has no buffer of its own.  In a forward transform it
    // uses this warp's rows of the X2 buffer of the PREVIOUS transform: those rows are read by this warp alone (forward X2
    // reads are partitioned by warp), the other warp writes that buffer again only one transform later, after the group
    // barrier this warp reaches with X1 long finished, and a forward transform never directly follows an inverse one without
    // the end-of-step barrier in between.  In an inverse transform X1 follows X2 and reuses the same buffer after an extra
    // barrier.  The 9 KiB per ciphertext this frees pay for a double-buffered X2: 10 group barriers per step instead of 16.
    static constexpr bool kX1Alias = ALIAS;
    // teams of two at KCT > 4 (168-register budget): both twiddle tables live in shared memory
    static constexpr bool kTwShared = TEAM == 2 && KCT > 4 && !ALIAS;
    static constexpr int kTw2Mode = kTwShared ? kTwSmem : (KCT > 4 ? kTwPow : kTwFull);   // KCT = 5, 6 without teams: keep r, r^2, r^4
    static constexpr int kTw3Mode = kTwShared ? kTwSmem : (KCT > 5 ? kTwPow : kTwFull);
    static constexpr bool kDbX2 = KCT <= 4 || ALIAS;     // double-buffered X2 (fits when only 4 groups share the SM, or without X1 buffers)
#ifdef TFHE_B200_STAGES
    static constexpr int kStages = KCT <= 4 ? TFHE_B200_STAGES : 3;   // variant builds: deeper ring where shared memory allows
#else
    static constexpr int kStages = 3;   // key-ring depth (4 measured no faster; KCT = 2 with a 2-deep ring and two CTAs, i.e. two
                                        // independent rings, per SM: 91.0 k/s, same as one KCT = 4 CTA -- ring coupling is not a limiter)
#endif
    static constexpr bool kAccTmem = !kTwShared && KCT > 5;   // MAC accumulators in TMEM (measured slower than KCT = 4)
    static constexpr int kTmemCols = 256;       // 64 columns per warp, up to 3 warps per TMEM quadrant
    static constexpr int kAccBytes = 2 * kN * 4;
    static constexpr int kX1Bytes = ALIAS ? 0 : kX1Slots * 16;
    static constexpr int kX2Bytes = kX2Slots * 16;
    static constexpr int kTwBytes = kTwShared ? (kTw2Len + kTw3Len) * 16 : 0;
    __host__ __device__ static constexpr int group_bytes(int n) {
        int b = kAccBytes + kX1Bytes + (kDbX2 ? 2 : 1) * kX2Bytes + align16((n + 1) * 2);
        // teams of two: consecutive groups sit 64 (mod 128) bytes apart, i.e. on complementary bank halves
        while (TEAM == 2 && (b & 127) != 64) b += 16;
        return b;
    }
};

// genDecompositionOffset (src/key.zig:121-131) of a gadget shape
__host__ __device__ constexpr uint32_t gadget_offset(int L, int bgbit) {
    uint32_t o = 0u;
    for (int i = 0; i < L; i++) o += (1u << (bgbit - 1)) << (32 - (i + 1) * bgbit);
    return o;
}
static_assert(gadget_offset(3, 6) == 0x82080000u && gadget_offset(1, 22) == 0x80000000u && gadget_offset(2, 10) == 0x80200000u, "key.zig:121-131");

template <int KCT, bool USE_TMA, bool MARGIN, int TEAM = 1, int LT = 0, int TWT = 0, int BGT = (LT == 3 ? 6 : LT == 1 ? 22 : 0)>
__global__ void __launch_bounds__(KCT * kGroupThreads, 1) blind_rotate_kernel(const BrArgs P) {
    static_assert((LT == 0) == (BGT == 0), "gadget length and digit width are compile-time constants together");
    using Lay = Layout<KCT, TEAM, (TWT != 0 && KCT > 4)>;
    static_assert(TWT == 0 || TEAM == 1 || KCT > 4, "teams of two with tensor-memory twiddles: six ciphertexts per CTA only");
    constexpr bool XA = Lay::kX1Alias;
    constexpr bool TWT_ON = TWT != 0;              // both twiddle sets in tensor memory, accumulators in registers
    constexpr int POW = TWT_ON ? kTwTmem : Lay::kTw2Mode, POW3 = TWT_ON ? kTwTmem : Lay::kTw3Mode;
    constexpr bool DBX2 = Lay::kDbX2;
    constexpr int kStages = Lay::kStages;
    constexpr bool ACCT = !TWT_ON && Lay::kAccTmem;
    constexpr bool USE_TMEM = ACCT || TWT_ON;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // ---- carve shared memory
    unsigned char *ptr = smem_raw;
    cplx *bsk_ring = reinterpret_cast<cplx *>(ptr);
    if (USE_TMA) ptr += kStages * kBskChunkBytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(ptr);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ptr + 64);
    uint32_t *ring_cnt = reinterpret_cast<uint32_t *>(ptr + 80);   // [kMaxStages] releases per stage, monotone
    ptr += 96;
    // LT / BGT > 0: gadget shape as compile-time constants.  3 / 6: the 80/110/128-bit sets; 1 / 22: UINT4 ... UINT8; 1 / 23: UINT3;
    // 1 / 18: UINT2; 2 / 10: UINT1 (all but the first with the wide rounding; the launcher checks offset and rounding mode)
    const int n = P.n, L = LT > 0 ? LT : P.L, bgbit = LT > 0 ? BGT : P.bgbit;
    static_assert(TEAM == 1 || (TEAM == 2 && KCT % 2 == 0), "a team never straddles CTAs");
    constexpr int kTeamThreads = TEAM * kGroupThreads;
    const int group_bytes = Lay::group_bytes(n);
    cplx *tw_tab = reinterpret_cast<cplx *>(ptr);   // [tw2 | tw3] when the twiddles live in shared memory
    ptr += Lay::kTwBytes;

    const int tid = threadIdx.x;
    const int first_ct = blockIdx.x * KCT;
    const int n_active = min(KCT, (int)P.B - first_ct);

    static_assert(Lay::kStages <= kMaxStages, "ring bookkeeping has room for kMaxStages stages");
    const uint32_t ring_warps = (uint32_t)(((n_active + TEAM - 1) / TEAM) * TEAM * 2);   // consumer warps of this CTA
    if (USE_TMA && tid == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], ring_warps);   // one arrival per consumer warp
            ring_cnt[s] = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (Lay::kTwShared) {
        for (int j = tid; j < kTw2Len + kTw3Len; j += KCT * kGroupThreads) tw_tab[j] = j < kTw2Len ? P.tw2[j] : P.tw3[j - kTw2Len];
    }
    if (USE_TMEM && tid < 32) {
        tmem_alloc(tmem_slot, Lay::kTmemCols);
        tmem_fence_before_sync();
    }
    __syncthreads();
    uint32_t tmem_base = 0;
    if (USE_TMEM) {
        tmem_fence_after_sync();
        tmem_base = *tmem_slot;
    }

    // TEAM = 2: two ciphertexts share each warp in adjacent lanes (lane = 2 j + c).  LDS.128 merges adjacent lane
    // pairs reading the same 16 bytes (tools/microbench.cu M1), so the key reads of the pointwise MAC -- identical
    // for both ciphertexts -- cost half the shared-memory wavefronts; the two ciphertexts' buffers differ by 64
    // (mod 128) bytes, so every quarter-warp phase of the exchanges still covers 8 distinct 16-byte bank groups.
    int g, t, team;
    if (TEAM == 1) {
        g = tid >> 6; t = tid & 63; team = g;
    } else {
        team = tid >> 7;
        t = 16 * ((tid >> 5) & 3) + ((tid & 31) >> 1);
        g = 2 * team + (tid & 1);
    }
    if (team * TEAM >= n_active) return;   // whole warps (warp 0 is always active: it frees the TMEM allocation at the end)
    const bool live = g < n_active;        // TEAM = 2: the ghost half of a last, odd team mirrors its partner
    const int hi = t >> 3, lo = t & 7;
    const int barid = 1 + team;
    // this warp's TMEM window: lane quadrant (warp & 3), 64 columns per warp sharing that quadrant
    const uint32_t my_tmem = tmem_base + ((uint32_t)((tid >> 5) & 3) << 21) + (uint32_t)(tid >> 7) * 64u;
    unsigned char *gb = ptr + (size_t)g * group_bytes;
    uint32_t *acc_a = reinterpret_cast<uint32_t *>(gb);
    uint32_t *acc_b = acc_a + kN;
    Xbuf xb;
    xb.x1 = reinterpret_cast<cplx *>(gb + Lay::kAccBytes);
    xb.x2 = reinterpret_cast<cplx *>(gb + Lay::kAccBytes + Lay::kX1Bytes);
    xb.flip = 0;
    uint16_t *atil = reinterpret_cast<uint16_t *>(gb + Lay::kAccBytes + Lay::kX1Bytes + (DBX2 ? 2 : 1) * Lay::kX2Bytes);
    const size_t ct = (size_t)P.ct_base + first_ct + (live ? g : g - 1);

    // per-thread twiddles: role B node q2 = lo, role C node t
    Tw2<POW> tw2;
    Tw2<POW3> tw3;
    if (TWT_ON) {   // park both sets in this warp's tensor-memory window: tw2 in columns [0, 32), tw3 in [32, 64)
        uint32_t r[32];
#pragma unroll
        for (int k = 28; k < 32; k++) r[k] = 0u;
#pragma unroll
        for (int p = 1; p < 8; p++) {
            const cplx w = P.tw2[tw2_index(p, lo)];
            r[4 * p - 4] = (uint32_t)__double2loint(w.re); r[4 * p - 3] = (uint32_t)__double2hiint(w.re);
            r[4 * p - 2] = (uint32_t)__double2loint(w.im); r[4 * p - 1] = (uint32_t)__double2hiint(w.im);
        }
        tmem_st32(my_tmem, r);
#pragma unroll
        for (int p = 1; p < 8; p++) {
            const cplx w = P.tw3[tw3_index(p, t)];
            r[4 * p - 4] = (uint32_t)__double2loint(w.re); r[4 * p - 3] = (uint32_t)__double2hiint(w.re);
            r[4 * p - 2] = (uint32_t)__double2loint(w.im); r[4 * p - 1] = (uint32_t)__double2hiint(w.im);
        }
        tmem_st32(my_tmem + 32u, r);
        tmem_wait_st();
        tw2.taddr = my_tmem;
        tw3.taddr = my_tmem + 32u;
    } else if (POW3 == kTwSmem) {
        tw3.tab = tw_tab + kTw2Len + tw3_index(1, t); tw3.stride = 64;
    } else if (POW3 == kTwPow) {
        tw3.w[0] = P.tw3[tw3_index(1, t)]; tw3.w[1] = P.tw3[tw3_index(2, t)]; tw3.w[2] = P.tw3[tw3_index(4, t)];
    } else {
#pragma unroll
        for (int p = 1; p < 8; p++) tw3.w[p - 1] = P.tw3[tw3_index(p, t)];
    }
    if (TWT_ON) {
    } else if (POW == kTwSmem) {
        tw2.tab = tw_tab + tw2_index(1, lo); tw2.stride = 8;
    } else if (POW == kTwPow) {
        tw2.w[0] = P.tw2[tw2_index(1, lo)]; tw2.w[1] = P.tw2[tw2_index(2, lo)]; tw2.w[2] = P.tw2[tw2_index(4, lo)];
    } else {
#pragma unroll
        for (int p = 1; p < 8; p++) tw2.w[p - 1] = P.tw2[tw2_index(p, lo)];
    }

    // ---- prologue: gate linear part (gates.zig:48-121) + modulus switch (trgsw.zig:297,312)
    {
        const GateOperands go = gate_operands(P, ct, n);
        const int op = go.op;
        for (int i = t; i <= n; i += kGroupThreads) {
            uint32_t lin = gate_linear_signed(go, i);
            if (i == n) lin += gate_constant(op);
            const uint32_t m = mod_switch_2n(lin, P.ms_shift);   // in [0, 2N]
            atil[i] = (uint16_t)((i == n) ? (2 * kN - m) : m);
        }
    }
    bar_sync(barid, kTeamThreads);
    // ---- acc = X^btil * testvec (trgsw.zig:300-306), stored in acc_pos order
    {
        const int btil = atil[n];
        const uint32_t *tv = P.testvec ? P.testvec + (P.tv_per_item ? ct * (size_t)(2 * kN) : 0) : nullptr;
        for (int j = t; j < kN; j += kGroupThreads) {
            const int u = (j - btil) & (2 * kN - 1);
            const uint32_t va = tv ? tv[u & (kN - 1)] : 0u;
            const uint32_t vb = tv ? tv[kN + (u & (kN - 1))] : 0x20000000u;   // key.zig:134-145
            acc_a[acc_pos(j)] = (u & kN) ? 0u - va : va;
            acc_b[acc_pos(j)] = (u & kN) ? 0u - vb : vb;
        }
    }
    bar_sync(barid, kTeamThreads);

    const uint32_t mask = (1u << bgbit) - 1u, half_bg = 1u << (bgbit - 1);
    constexpr uint32_t kOffset = gadget_offset(LT > 0 ? LT : 1, BGT > 0 ? BGT : 1);
    const uint32_t offset = LT > 0 ? kOffset : P.offset;   // the launcher checks it
    const int wide = LT == 3 ? 0 : LT > 0 ? 1 : P.wide_round;   // BGBIT = 6: magic-add rounding (coefficients < 2^45)
    int stage = 0;
    uint32_t phase = 0;
    double margin = 0.0;
    Producer pr;
    pr.ring = bsk_ring; pr.full_bar = full_bar; pr.empty_bar = empty_bar; pr.stages = kStages;
    pr.ring_cnt = ring_cnt; pr.bsk = P.bsk; pr.total = (uint32_t)(n * 2 * L); pr.ring_warps = ring_warps;
#ifdef TFHE_B200_DIAG
    const int diag = P.diag;
#else
    constexpr int diag = 0;
#endif
    if (kRingLastArriver) {
        pr.active = false; pr.remaining = 0;
        if (USE_TMA && tid == 0 && !DIAG_ON(P, 0)) {   // fill the ring; every later bulk copy is issued by the consumers (ring_check)
            const uint64_t policy = l2_policy_evict_last();
            for (int s = 0; s < kStages && s < n * 2 * L; s++) {
                mbar_arrive_expect_tx(&full_bar[s], kBskChunkBytes);
                bulk_g2s(bsk_ring + s * kBskChunkCplx, P.bsk + (size_t)s * kBskChunkCplx, kBskChunkBytes, &full_bar[s], policy);
            }
        }
    } else {
        pr.src = P.bsk;
        pr.remaining = (USE_TMA && !DIAG_ON(P, 0)) ? n * 2 * L : 0; pr.issued = 0; pr.stage = 0; pr.phase = 0;
        pr.active = USE_TMA && tid == 0;
        pr.policy = pr.active ? l2_policy_evict_last() : 0;
        if (USE_TMA) {   // fill the ring
#pragma unroll
            for (int s = 0; s < kStages; s++) producer_poll(pr);
        }
    }

    // CTAs of one wave start in lock step and would ask L2 for the same 16 KiB key chunk in the same few hundred cycles for
    // the whole launch (measured: a single wave takes 6.5 ms, the 70th of a long launch 6.1 ms, once the CTAs have drifted
    // apart).  A few microseconds of per-CTA stagger at the start spreads them from the first step on.
    if (USE_TMA) __nanosleep((blockIdx.x % 41u) * 128u);
    // ---- n CMUX steps (trgsw.zig:311-330)
    for (int i = 0; i < n; i++) {
        const int at = atil[i];
        if (!ACCT) {
            // ---- accumulators in registers (252-register budget)
            cplx oa[8], ob[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { oa[q] = cplx{0.0, 0.0}; ob[q] = cplx{0.0, 0.0}; }
// (unrolled at L = 3: 97.2 k instead of 103.0 k bootstraps/s at six per CTA; at L = 1, one digit transform per polynomial, it pays: UINT4
// fast mode 187.6 k -> 191.9 k, UINT2 223.4 k -> 228.6 k blind rotations/s, as in the exact kernel)
#pragma unroll (LT == 1 ? 2 : 1)
            for (int h = 0; h < 2; h++) {          // h = 0: digits of the a polynomial, 1: of b (trgsw.zig:211-217)
                const uint32_t *accp = h ? acc_b : acc_a;
                uint32_t d[16];
                load_rot_diffs(d, accp, at, offset, hi, lo);
#pragma unroll (LT > 0 ? LT : 1)   // (rolled, constants still compile-time: 96.5 k instead of 103.0 k bootstraps/s at six per CTA)
                for (int l = 0; l < L; l++) {
                    cplx v[8];
                    digits_to_cplx(v, d, 32 - (l + 1) * bgbit, mask, half_bg);
                    fwd_transform<USE_TMA, DBX2, POW, POW3, XA>(v, xb, tw2, tw3, hi, lo, barid, pr, kTeamThreads, diag);
                    const cplx *chunk;
                    if (USE_TMA) {
                        if (!DIAG_ON(P, 0) && !DIAG_ON(P, 5)) {
                            if (kRingLastArriver) mbar_wait(&full_bar[stage], phase);
                            else
                                while (!mbar_try_wait(&full_bar[stage], phase)) producer_poll(pr);
                        }
                        chunk = bsk_ring + stage * kBskChunkCplx;
                    } else {
                        chunk = P.bsk + ((size_t)i * 2 * L + h * L + l) * kBskChunkCplx;
                    }
                    if (DIAG_ON(P, 1)) {
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            cmac(oa[q], v[q], cplx{1.0, 0.5});
                            cmac(ob[q], v[q], cplx{0.25, 1.0});
                        }
                    } else {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        cmac(oa[q], v[q], chunk[bsk_slot(0, q, t)]);
                        cmac(ob[q], v[q], chunk[bsk_slot(1, q, t)]);
                    }
                    }
                    if (USE_TMA && !DIAG_ON(P, 0)) {
                        if (kRingLastArriver) ring_release(pr, stage, (uint32_t)((i * 2 + h) * L + l), tid & 31);
                        else {
                            __syncwarp();
                            if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if constexpr (XA && DBX2 && kPairInverse) {
                inv_transform_pair<POW, POW3>(oa, ob, xb, tw2, tw3, hi, lo, barid, kTeamThreads, [&](const cplx (&v)[8], int which) {
                    round_accumulate<MARGIN>(v, which ? acc_b : acc_a, t, wide, margin);
                });
            } else {
                inv_transform<USE_TMA, DBX2, POW, POW3, XA>(oa, xb, tw2, tw3, hi, lo, barid, pr, kTeamThreads, diag);
                round_accumulate<MARGIN>(oa, acc_a, t, wide, margin);
                inv_transform<USE_TMA, DBX2, POW, POW3, XA>(ob, xb, tw2, tw3, hi, lo, barid, pr, kTeamThreads, diag);
                round_accumulate<MARGIN>(ob, acc_b, t, wide, margin);
            }
        } else {
            // ---- accumulators in TMEM (168-register budget): each half is loaded, updated and stored back
#pragma unroll 1
            for (int h = 0; h < 2; h++) {
                const uint32_t *accp = h ? acc_b : acc_a;
                uint32_t d[16];
                load_rot_diffs(d, accp, at, offset, hi, lo);
#pragma unroll 1
                for (int l = 0; l < L; l++) {
                    cplx v[8];
                    digits_to_cplx(v, d, 32 - (l + 1) * bgbit, mask, half_bg);
                    fwd_transform<USE_TMA, DBX2, POW, POW3>(v, xb, tw2, tw3, hi, lo, barid, pr, kTeamThreads);
                    const cplx *chunk;
                    if (USE_TMA) {
                        if (kRingLastArriver) mbar_wait(&full_bar[stage], phase);
                        else
                            while (!mbar_try_wait(&full_bar[stage], phase)) producer_poll(pr);
                        chunk = bsk_ring + stage * kBskChunkCplx;
                    } else {
                        chunk = P.bsk + ((size_t)i * 2 * L + h * L + l) * kBskChunkCplx;
                    }
                    const bool first = (h == 0 && l == 0);
#pragma unroll 1
                    for (int ab = 0; ab < 2; ab++) {
                        cplx o[8];
                        if (first) {
#pragma unroll
                            for (int q = 0; q < 8; q++) o[q] = cplx{0.0, 0.0};
                        } else {
                            tmem_load_cplx8(o, my_tmem + 32 * ab);
                        }
#pragma unroll
                        for (int q = 0; q < 8; q++) cmac(o[q], v[q], chunk[bsk_slot(ab, q, t)]);
                        tmem_store_cplx8(my_tmem + 32 * ab, o);
                    }
                    if (USE_TMA) {
                        if (kRingLastArriver) ring_release(pr, stage, (uint32_t)((i * 2 + h) * L + l), tid & 31);
                        else {
                            __syncwarp();
                            if ((tid & 31) == 0) mbar_arrive(&empty_bar[stage]);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
#pragma unroll 1
            for (int ab = 0; ab < 2; ab++) {
                cplx o[8];
                tmem_load_cplx8(o, my_tmem + 32 * ab);
                inv_transform<USE_TMA, DBX2, POW, POW3>(o, xb, tw2, tw3, hi, lo, barid, pr, kTeamThreads);
                round_accumulate<MARGIN>(o, ab ? acc_b : acc_a, t, wide, margin);
            }
        }
        // accumulator complete before the next rotated reads.  With a double-buffered X2 no barrier is needed here: the a half's
        // add-back is followed by a group barrier of the b half's inverse transform (the pair's third barrier), the b half is not read
        // before at least one more group barrier, of the next step's first digit transform, and the buffers the next step touches
        // first are free by then (tests/test_exchange_protocol.py checks this protocol as a model, for every kernel layout).
        if constexpr (!DBX2 || ACCT) bar_sync(barid, kTeamThreads);
    }
    if constexpr (DBX2 && !ACCT) bar_sync(barid, kTeamThreads);   // the epilogue reads other threads' coefficients

    // ---- epilogue
    if (live && P.out_trlwe) {
        uint32_t *o = P.out_trlwe + ct * (size_t)(2 * kN);
        for (int j = t; j < kN; j += kGroupThreads) { o[j] = acc_a[acc_pos(j)]; o[kN + j] = acc_b[acc_pos(j)]; }
    }
    if (live && P.out_lv1) {   // sampleExtractIndex(., 0): trlwe.zig:146-162
        uint32_t *o = P.out_lv1 + ct * (size_t)(kN + 1);
        for (int j = t; j <= kN; j += kGroupThreads)
            o[j] = (j == 0) ? acc_a[0] : (j == kN) ? acc_b[0] : 0u - acc_a[acc_pos(kN - j)];
    }
    if (MARGIN && P.margin_bits) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) margin = fmax(margin, __shfl_xor_sync(0xffffffffu, margin, s));
        if ((tid & 31) == 0) atomicMax(P.margin_bits, (unsigned long long)__double_as_longlong(margin));
    }
    if (USE_TMEM) {
        if (ACCT) tmem_wait_st();
        tmem_fence_before_sync();
        bar_sync(15, ((n_active + TEAM - 1) / TEAM) * kTeamThreads);      // every active warp is done with its TMEM window
        if (tid < 32) {
            tmem_fence_after_sync();
            tmem_dealloc(tmem_base, Lay::kTmemCols);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Latency mode (batches no larger than the SM count): one CTA per ciphertext, the 2L digit transforms of an
// iteration run concurrently on 2L groups of 64 threads instead of one after another.  Group r transforms
// digit polynomial r and multiplies it with key row r (loaded straight from L2 at the top of the iteration);
// the 2L products are reduced through shared memory by groups 0 (a part) and 1 (b part), which then run the
// two inverse transforms.  Same building blocks and layouts as the throughput kernel; the sum over rows is
// associated differently (separately rounded products), which rounds to the same integers wherever the
// external product is exact (the launcher only selects this kernel for those parameter sets).
template <int L, bool MARGIN>
__global__ void __launch_bounds__(2 * L * kGroupThreads, 1) blind_rotate_latency_kernel(const BrArgs P) {
    constexpr int G = 2 * L;
    constexpr int POW = (L >= 3) ? kTwPow : kTwFull;    // 384 threads -> 168 registers: expand twiddle powers per pass
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *acc_a = reinterpret_cast<uint32_t *>(smem_raw), *acc_b = acc_a + kN;
    cplx *red = reinterpret_cast<cplx *>(smem_raw + 2 * kN * 4);                      // [G][ab][q][t]
    unsigned char *xbase = smem_raw + 2 * kN * 4 + (size_t)G * kBskChunkBytes;
    constexpr int kXBytes = (kX1Slots + kX2Slots) * 16;
    uint64_t *key_bar = reinterpret_cast<uint64_t *>(xbase + G * kXBytes);           // one mbarrier per group
    uint16_t *atil = reinterpret_cast<uint16_t *>(xbase + G * kXBytes + 64);
    const int n = P.n, bgbit = P.bgbit;
    const int tid = threadIdx.x, g = tid >> 6, t = tid & 63, hi = t >> 3, lo = t & 7;
    const int barid = 1 + g;
    const size_t ct = (size_t)P.ct_base + blockIdx.x;
    // Key row r of iteration i lands by cp.async.bulk directly in this group's slice of the reduction buffer
    // (identical [ab][q][t] layout): the thread that reads a key value overwrites it with its product.
    cplx *my_red = red + (size_t)g * kBskChunkCplx;
    const uint64_t policy = l2_policy_evict_last();
    if (t == 0) {
        mbar_init(&key_bar[g], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(&key_bar[g], kBskChunkBytes);
        bulk_g2s(my_red, P.bsk + (size_t)g * kBskChunkCplx, kBskChunkBytes, &key_bar[g], policy);
    }
    Xbuf xb;
    xb.x1 = reinterpret_cast<cplx *>(xbase + g * kXBytes);
    xb.x2 = xb.x1 + kX1Slots;
    xb.flip = 0;
    Tw2<POW> tw2;
    Tw2<POW> tw3;
    if (POW == kTwPow) {
        tw2.w[0] = P.tw2[tw2_index(1, lo)]; tw2.w[1] = P.tw2[tw2_index(2, lo)]; tw2.w[2] = P.tw2[tw2_index(4, lo)];
        tw3.w[0] = P.tw3[tw3_index(1, t)]; tw3.w[1] = P.tw3[tw3_index(2, t)]; tw3.w[2] = P.tw3[tw3_index(4, t)];
    } else {
#pragma unroll
        for (int p = 1; p < 8; p++) { tw2.w[p - 1] = P.tw2[tw2_index(p, lo)]; tw3.w[p - 1] = P.tw3[tw3_index(p, t)]; }
    }
    {   // prologue: gate linear part + modulus switch, whole CTA
        const GateOperands go = gate_operands(P, ct, n);
        const int op = go.op;
        for (int i = tid; i <= n; i += G * kGroupThreads) {
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
        for (int j = tid; j < kN; j += G * kGroupThreads) {
            const int u = (j - btil) & (2 * kN - 1);
            const uint32_t va = tv ? tv[u & (kN - 1)] : 0u;
            const uint32_t vb = tv ? tv[kN + (u & (kN - 1))] : 0x20000000u;
            acc_a[acc_pos(j)] = (u & kN) ? 0u - va : va;
            acc_b[acc_pos(j)] = (u & kN) ? 0u - vb : vb;
        }
    }
    __syncthreads();

    const uint32_t mask = (1u << bgbit) - 1u, half_bg = 1u << (bgbit - 1);
    const int wide = P.wide_round;
    const int h = g / L, l = g - h * L;
    const int sh = 32 - (l + 1) * bgbit;
    double margin = 0.0;
    Producer pr;
    pr.active = false; pr.remaining = 0;
    for (int i = 0; i < n; i++) {
        const int at = atil[i];
        cplx v[8];
        {
            uint32_t d[16];
            load_rot_diffs(d, h ? acc_b : acc_a, at, P.offset, hi, lo);
            digits_to_cplx(v, d, sh, mask, half_bg);
        }
        fwd_transform<false, false, POW, POW>(v, xb, tw2, tw3, hi, lo, barid, pr);
        mbar_wait(&key_bar[g], (uint32_t)(i & 1));
#pragma unroll
        for (int q = 0; q < 8; q++) {
            my_red[bsk_slot(0, q, t)] = cmul(v[q], my_red[bsk_slot(0, q, t)]);
            my_red[bsk_slot(1, q, t)] = cmul(v[q], my_red[bsk_slot(1, q, t)]);
        }
        __syncthreads();
        if (g < 2) {   // group 0 reduces and inverts the a part, group 1 the b part
            cplx o[8];
#pragma unroll
            for (int q = 0; q < 8; q++) o[q] = red[g * 512 + q * 64 + t];
#pragma unroll
            for (int r = 1; r < G; r++) {
#pragma unroll
                for (int q = 0; q < 8; q++) o[q] = cadd(o[q], red[(r * 2 + g) * 512 + q * 64 + t]);
            }
            // all 2L products are consumed (both reducers read every group's slice): the barrier below orders those
            // generic-proxy reads before the async-proxy refill issued after it
            inv_transform<false, false, POW, POW>(o, xb, tw2, tw3, hi, lo, barid, pr);
            round_accumulate<MARGIN>(o, g ? acc_b : acc_a, t, wide, margin);
        }
        __syncthreads();
        if (t == 0 && i + 1 < n) {   // prefetch the next iteration's key row; its latency hides behind the forward transform
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(&key_bar[g], kBskChunkBytes);
            bulk_g2s(my_red, P.bsk + ((size_t)(i + 1) * G + g) * kBskChunkCplx, kBskChunkBytes, &key_bar[g], policy);
        }
    }
    if (P.out_trlwe) {
        uint32_t *o = P.out_trlwe + ct * (size_t)(2 * kN);
        for (int j = tid; j < kN; j += G * kGroupThreads) { o[j] = acc_a[acc_pos(j)]; o[kN + j] = acc_b[acc_pos(j)]; }
    }
    if (P.out_lv1) {
        uint32_t *o = P.out_lv1 + ct * (size_t)(kN + 1);
        for (int j = tid; j <= kN; j += G * kGroupThreads)
            o[j] = (j == 0) ? acc_a[0] : (j == kN) ? acc_b[0] : 0u - acc_a[acc_pos(kN - j)];
    }
    if (MARGIN && P.margin_bits) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) margin = fmax(margin, __shfl_xor_sync(0xffffffffu, margin, s));
        if ((t & 31) == 0) atomicMax(P.margin_bits, (unsigned long long)__double_as_longlong(margin));
    }
}

// (Tangent-form twiddles -- w = c (1 + i t) with the scale c folded into the first butterfly layer, 72 instead of 80 FP64
// operations per forward pass -- were built and measured as well: bit-identical results, 242 registers, 89.6 k instead of
// 90.9 k bootstraps/s at KCT = 4.  Fewer FP64 instructions do not shorten the dependent chain a warp waits on.)
// (A software-pipelined variant of the throughput kernel -- X2 exchanges guarded by split mbarrier arrive / wait so that
// pass 1 of the next digit transform ran between the stores and the loads of the current one -- was built and measured:
// bit-identical results, 84.6 k instead of 90.8 k bootstraps/s.  mbarrier try_wait polling costs more than bar.sync and the
// exposed latency is that of the loads themselves, not of the barrier; see DESIGN.md section 4 and
// profiles/r01_team_probe.log.  The code was removed again.)

// ---------------------------------------------------------------------------------------------------
// Latency mode on a pair of SMs (batches of at most sm_count / 2 ciphertexts): a thread-block cluster of two CTAs per
// ciphertext.  CTA h owns polynomial h of the accumulator (h = 0: a, 1: b): its L groups of 64 threads transform the L
// digit polynomials of that half concurrently and multiply them with their key rows (both output parts); the L
// products per output are summed inside the CTA; the partial sum for the OTHER CTA's output travels through
// distributed shared memory (8 KiB per CTA and step) as asynchronous stores that complete a transaction count on an
// mbarrier of the receiving CTA (st.async ... mbarrier::complete_tx: no release fence, no cluster barrier inside the
// loop -- the first version used barrier.cluster per step and spent 15 % of its time in MEMBAR.ALL.GPU / ERRBAR),
// and each CTA runs the inverse transform of its own output and updates its own accumulator half -- the
// only state a step of blind rotation needs from the other half is that partial sum.  Against the single-CTA latency
// kernel: half the transforms per SM (FP64 pipe: 3 + 1 instead of 6 + 2 transforms per step on one SM), 192 threads per
// CTA so every twiddle is register-resident (no power expansion on the critical path), and the two inverse transforms
// on different SMs.  Same exactness precondition as the single-CTA latency kernel (a different summation order).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_cplx(uint32_t addr, cplx v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.re), "d"(v.im) : "memory");
}
// asynchronous remote store that reports its 16 bytes to an mbarrier of the destination CTA: the receiver waits on that
// barrier's transaction count, no release fence or cluster barrier on either side
__device__ __forceinline__ void st_async_cplx(uint32_t addr, cplx v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(addr), "d"(v.re), "d"(v.im),
                 "r"(remote_mbar)
                 : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int L>
__global__ void __launch_bounds__(L * kGroupThreads, 1) blind_rotate_pair_kernel(const BrArgs P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *acc = reinterpret_cast<uint32_t *>(smem_raw);                                 // this CTA's polynomial, acc_pos order
    cplx *red2 = reinterpret_cast<cplx *>(smem_raw + kN * 4);                                // [2 (step parity)][L][ab][q][t]: key rows, then products
    cplx *recv = red2 + (size_t)2 * L * kBskChunkCplx;                                       // [2 (step parity)][512]: partner's partial sum
    unsigned char *xbase = reinterpret_cast<unsigned char *>(recv + 2 * kHalfN);
    constexpr int kXBytes = (kX1Slots + kX2Slots) * 16;
    uint64_t *key_bar = reinterpret_cast<uint64_t *>(xbase + L * kXBytes);                   // [2 (step parity)][L] mbarriers
    uint64_t *recv_bar = key_bar + 2 * L;                                                    // [2 (step parity)]: partner's partial sum landed
    uint16_t *atil = reinterpret_cast<uint16_t *>(xbase + L * kXBytes + 64);
    // Key rows are double-buffered: the row of step i + 1 is requested at the top of step i (its buffer held the products
    // of step i - 1, all consumed before that step's cluster barrier), so the bulk copy has a whole step to land.
    const int n = P.n, bgbit = L == 3 ? 6 : P.bgbit;   // the launcher routes L = 3 here only with BGBIT = 6
    const int tid = threadIdx.x, g = tid >> 6, t = tid & 63, hi = t >> 3, lo = t & 7;
    const int barid = 1 + g;
    const uint32_t h = cluster_ctarank();            // which polynomial of the accumulator this CTA owns
    const size_t ct = (size_t)P.ct_base + (blockIdx.x >> 1);
    const int r = (int)h * L + g;                    // gadget row of this group (trgsw.zig:211-217)
    const uint64_t policy = l2_policy_evict_last();
    if (t == 0) {
        mbar_init(&key_bar[g], 1);
        mbar_init(&key_bar[L + g], 1);
        if (g == 0) { mbar_init(&recv_bar[0], 1); mbar_init(&recv_bar[1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(&key_bar[g], kBskChunkBytes);
        bulk_g2s(red2 + (size_t)g * kBskChunkCplx, P.bsk + (size_t)r * kBskChunkCplx, kBskChunkBytes, &key_bar[g], policy);
    }
    Xbuf xb;
    xb.x1 = reinterpret_cast<cplx *>(xbase + g * kXBytes);
    xb.x2 = xb.x1 + kX1Slots;
    xb.flip = 0;
    Tw2<kTwFull> tw2, tw3;
#pragma unroll
    for (int p = 1; p < 8; p++) { tw2.w[p - 1] = P.tw2[tw2_index(p, lo)]; tw3.w[p - 1] = P.tw3[tw3_index(p, t)]; }
    {   // prologue: gate linear part + modulus switch (both CTAs need every rotation amount)
        const GateOperands go = gate_operands(P, ct, n);
        const int op = go.op;
        for (int i = tid; i <= n; i += L * kGroupThreads) {
            uint32_t lin = gate_linear_signed(go, i);
            if (i == n) lin += gate_constant(op);
            const uint32_t m = mod_switch_2n(lin, P.ms_shift);
            atil[i] = (uint16_t)((i == n) ? (2 * kN - m) : m);
        }
    }
    __syncthreads();
    {
        const int btil = atil[n];
        const uint32_t *tv = P.testvec ? P.testvec + (P.tv_per_item ? ct * (size_t)(2 * kN) : 0) + h * kN : nullptr;
        for (int j = tid; j < kN; j += L * kGroupThreads) {
            const int u = (j - btil) & (2 * kN - 1);
            const uint32_t v = tv ? tv[u & (kN - 1)] : (h ? 0x20000000u : 0u);   // key.zig:134-145
            acc[acc_pos(j)] = (u & kN) ? 0u - v : v;
        }
    }
    // the partner's receive buffer, seen from here
    const uint32_t remote_recv = map_to_cta(smem_u32(recv), h ^ 1u);
    const uint32_t remote_bar = map_to_cta(smem_u32(recv_bar), h ^ 1u);
    cluster_arrive();     // both CTAs of the pair are resident and initialised before any remote store
    cluster_wait();

    const uint32_t mask = (1u << bgbit) - 1u, half_bg = 1u << (bgbit - 1);
    const int sh = 32 - (g + 1) * bgbit;
    const int wide = L == 3 ? 0 : P.wide_round;
    double margin = 0.0;
    Producer pr;
    pr.active = false; pr.remaining = 0;
    for (int i = 0; i < n; i++) {
        const int at = atil[i];
        cplx *red = red2 + (size_t)(i & 1) * L * kBskChunkCplx;
        cplx *my_red = red + (size_t)g * kBskChunkCplx;
        if (t == 0 && i + 1 < n) {   // request the next step's key row into the other buffer
            const int nb = (i + 1) & 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(&key_bar[nb * L + g], kBskChunkBytes);
            bulk_g2s(red2 + ((size_t)nb * L + g) * kBskChunkCplx, P.bsk + ((size_t)(i + 1) * 2 * L + r) * kBskChunkCplx, kBskChunkBytes,
                     &key_bar[nb * L + g], policy);
        }
        cplx v[8];
        {
            uint32_t d[16];
            load_rot_diffs(d, acc, at, P.offset, hi, lo);
            digits_to_cplx(v, d, sh, mask, half_bg);
        }
        fwd_transform<false, false, kTwFull, kTwFull>(v, xb, tw2, tw3, hi, lo, barid, pr);
        mbar_wait(&key_bar[(i & 1) * L + g], (uint32_t)((i >> 1) & 1));
        // products for the OTHER CTA's output first: their sum is on its way through distributed shared memory (the slowest
        // hop of the step) while every group multiplies for this CTA's own output
        const uint32_t other = h ^ 1u;
#pragma unroll
        for (int q = 0; q < 8; q++) my_red[bsk_slot(other, q, t)] = cmul(v[q], my_red[bsk_slot(other, q, t)]);
        if (tid == 0) mbar_arrive_expect_tx(&recv_bar[i & 1], kHalfN * 16);   // this step's 8 KiB from the partner
        __syncthreads();
        cplx o[8];
        // the partial sum of the OTHER output goes to the partner (group L - 1), our own stays in registers (group 0)
        if (g == L - 1) {
#pragma unroll
            for (int q = 0; q < 8; q++) o[q] = red[other * 512 + q * 64 + t];
#pragma unroll
            for (int l = 1; l < L; l++) {
#pragma unroll
                for (int q = 0; q < 8; q++) o[q] = cadd(o[q], red[(l * 2 + other) * 512 + q * 64 + t]);
            }
            const uint32_t dst = remote_recv + (uint32_t)(((i & 1) * kHalfN + t) * 16);
#pragma unroll
            for (int q = 0; q < 8; q++) st_async_cplx(dst + q * 64 * 16, o[q], remote_bar + (uint32_t)((i & 1) * 8));
        }
#pragma unroll
        for (int q = 0; q < 8; q++) my_red[bsk_slot(h, q, t)] = cmul(v[q], my_red[bsk_slot(h, q, t)]);
        __syncthreads();
        if (g == 0) {
#pragma unroll
            for (int q = 0; q < 8; q++) o[q] = red[h * 512 + q * 64 + t];
#pragma unroll
            for (int l = 1; l < L; l++) {
#pragma unroll
                for (int q = 0; q < 8; q++) o[q] = cadd(o[q], red[(l * 2 + h) * 512 + q * 64 + t]);
            }
        }
        if (g == 0) {
            mbar_wait(&recv_bar[i & 1], (uint32_t)((i >> 1) & 1));   // the partner's 512 asynchronous stores have completed
            const cplx *rv = recv + (i & 1) * kHalfN;
#pragma unroll
            for (int q = 0; q < 8; q++) o[q] = cadd(o[q], rv[q * 64 + t]);
            inv_transform<false, false, kTwFull, kTwFull>(o, xb, tw2, tw3, hi, lo, barid, pr);
            round_accumulate<false>(o, acc, t, wide, margin);
        }
        __syncthreads();      // accumulator half complete before the next rotated reads
    }
    if (P.out_trlwe) {
        uint32_t *o = P.out_trlwe + ct * (size_t)(2 * kN) + h * kN;
        for (int j = tid; j < kN; j += L * kGroupThreads) o[j] = acc[acc_pos(j)];
    }
    if (P.out_lv1) {   // sampleExtractIndex(., 0): trlwe.zig:146-162 -- mask from the a half, body from the b half
        uint32_t *o = P.out_lv1 + ct * (size_t)(kN + 1);
        if (h == 0) {
            for (int j = tid; j < kN; j += L * kGroupThreads) o[j] = (j == 0) ? acc[0] : 0u - acc[acc_pos(kN - j)];
        } else if (tid == 0) {
            o[kN] = acc[0];
        }
    }
    cluster_arrive();     // neither CTA exits while its shared memory may still be written by the other
    cluster_wait();
}

template <int L>
cudaError_t launch_pair(const BrArgs &a, cudaStream_t s) {
    const size_t smem = kN * 4 + (size_t)2 * L * kBskChunkBytes + 2 * kHalfN * 16 + (size_t)L * (kX1Slots + kX2Slots) * 16 + 64 + align16((a.n + 1) * 2);
    auto kern = blind_rotate_pair_kernel<L>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * a.B);
    cfg.blockDim = dim3(L * kGroupThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

template <int L, bool MARGIN>
cudaError_t launch_latency(const BrArgs &a, cudaStream_t s) {
    constexpr int G = 2 * L;
    const size_t smem = 2 * kN * 4 + (size_t)G * kBskChunkBytes + (size_t)G * (kX1Slots + kX2Slots) * 16 + 64 + align16((a.n + 1) * 2);
    auto kern = blind_rotate_latency_kernel<L, MARGIN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<a.B, G * kGroupThreads, smem, s>>>(a);
    return cudaGetLastError();
}

// tensor-memory twiddles (tuning key "twt"): throughput kernel at KCT = 4, 5, 6, TMA ring, no margin tracking
template <int KCT, int TEAM = 1>
cudaError_t launch_twt(const BrArgs &a, cudaStream_t s) {
    using Lay = Layout<KCT, TEAM, (KCT > 4)>;
    const size_t smem = Lay::kStages * kBskChunkBytes + 96 + (size_t)KCT * Lay::group_bytes(a.n);
    auto kern = blind_rotate_kernel<KCT, true, false, TEAM, 0, 1>;
    if (a.L == 3 && a.bgbit == 6 && a.offset == 0x82080000u && !a.wide_round && kUnrollL3) kern = blind_rotate_kernel<KCT, true, false, TEAM, 3, 1>;
    constexpr bool kInst = KCT == 6;   // the UINT gadget shapes: at the full-wave width only
    if (kInst && a.wide_round && kUnrollL3 && a.L >= 1 && a.L <= 2 && a.bgbit >= 1 && a.L * a.bgbit <= 32 && a.offset == gadget_offset(a.L, a.bgbit)) {
        if (a.L == 1 && a.bgbit == 22) kern = blind_rotate_kernel<KCT, true, false, TEAM, (kInst ? 1 : 0), 1, (kInst ? 22 : 0)>;        // UINT4 ... UINT8
        else if (a.L == 1 && a.bgbit == 23) kern = blind_rotate_kernel<KCT, true, false, TEAM, (kInst ? 1 : 0), 1, (kInst ? 23 : 0)>;   // UINT3
        else if (a.L == 1 && a.bgbit == 18) kern = blind_rotate_kernel<KCT, true, false, TEAM, (kInst ? 1 : 0), 1, (kInst ? 18 : 0)>;   // UINT2
        else if (a.L == 2 && a.bgbit == 10) kern = blind_rotate_kernel<KCT, true, false, TEAM, (kInst ? 2 : 0), 1, (kInst ? 10 : 0)>;   // UINT1
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(a.B + KCT - 1) / KCT, KCT * kGroupThreads, smem, s>>>(a);
    return cudaGetLastError();
}

template <int KCT, bool USE_TMA, bool MARGIN, int TEAM = 1>
cudaError_t launch_variant(const BrArgs &a, cudaStream_t s) {
    using Lay = Layout<KCT, TEAM>;
    const int threads = KCT * kGroupThreads;
    const size_t smem = (USE_TMA ? Lay::kStages * kBskChunkBytes : 0) + 96 + Lay::kTwBytes + (size_t)KCT * Lay::group_bytes(a.n);
    auto kern = blind_rotate_kernel<KCT, USE_TMA, MARGIN, TEAM>;
    // the L = 3 / BGBIT = 6 sets (80/110/128-bit) get their own instantiation: digit loop unrolled, shifts, masks and the
    // rounding mode compile-time (+5 % on the 128-bit bench; profiles/r01_wave_scaling.log)
    if (KCT <= 4 && USE_TMA && !MARGIN && a.L == 3 && a.bgbit == 6 && a.offset == 0x82080000u && !a.wide_round && kUnrollL3)
        kern = blind_rotate_kernel<KCT, USE_TMA, MARGIN, TEAM, (KCT <= 4 && USE_TMA && !MARGIN) ? 3 : 0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (a.B + KCT - 1) / KCT;
    kern<<<grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}

template <int KCT>
cudaError_t launch_team2(const BrArgs &a, bool margin, cudaStream_t s) {
    return margin ? launch_variant<KCT, true, true, 2>(a, s) : launch_variant<KCT, true, false, 2>(a, s);
}

template <int KCT>
cudaError_t launch_kct(const BrArgs &a, bool tma, bool margin, cudaStream_t s) {
    if (tma) return margin ? launch_variant<KCT, true, true>(a, s) : launch_variant<KCT, true, false>(a, s);
    return margin ? launch_variant<KCT, false, true>(a, s) : launch_variant<KCT, false, false>(a, s);
}

}  // namespace

cudaError_t launch_blind_rotate(const BrArgs &a, const BrTuning &tune, bool track_margin, cudaStream_t s, uint64_t *launches) {
    if (a.B == 0) return cudaSuccess;
    const unsigned sm_total = tune.sm_count > 0 ? (unsigned)tune.sm_count : 148u;
    // latency mode: every ciphertext gets an SM of its own; only where the external product is exact
    // (so the different summation order cannot change a rounded coefficient)
    if (tune.latency_mode == 1 && tune.kct <= 0 && 2 * a.B <= sm_total && !a.wide_round && !track_margin && a.L >= 1 && a.L <= 3 && (a.L != 3 || a.bgbit == 6)) {
        if (launches) (*launches)++;     // one ciphertext per pair of SMs
        switch (a.L) {
            case 1: return launch_pair<1>(a, s);
            case 2: return launch_pair<2>(a, s);
            default: return launch_pair<3>(a, s);
        }
    }
    if (tune.latency_mode != 0 && tune.kct <= 0 && a.B <= sm_total && !a.wide_round && a.L >= 1 && a.L <= 3) {
        if (launches) (*launches)++;
        switch (a.L) {
            case 1: return track_margin ? launch_latency<1, true>(a, s) : launch_latency<1, false>(a, s);
            case 2: return track_margin ? launch_latency<2, true>(a, s) : launch_latency<2, false>(a, s);
            default: return track_margin ? launch_latency<3, true>(a, s) : launch_latency<3, false>(a, s);
        }
    }
    // CTA widths and the time one CTA of that width takes, measured on B200 at n = 700 (profiles/r01_wave_scaling.log,
    // profiles/r02_k1_ring.log): 1 or 2 ciphertexts 4.6 ms, 3: 5.8 ms, 4: 6.15 ms (96.3 k bootstraps/s), 6 with the twiddles in
    // tensor memory and X1 laid over the spare X2 buffer: 8.62 ms (103.1 k/s) -- the densest configuration wherever it is available.
    const bool twt_ok = tune.twt >= 0 && tune.use_tma != 0 && !track_margin && tune.team != 2;
    static const int widths[5] = {1, 2, 3, 4, 6};
    static const double t_cta[5] = {4.6, 4.6, 5.8, 6.15, 8.62};
    const int n_widths = twt_ok ? 5 : 4;
    const int dense = widths[n_widths - 1];
    const double t_dense = t_cta[n_widths - 1];
    int kct = tune.kct;
    if (kct <= 0 && tune.concurrent != 0) kct = dense;   // densest CTA (ciphertexts per SM-second); idle SMs go to the other lanes
    if (kct <= 0 && tune.use_tma != 0 && a.ct_base == 0) {
        // Mid-size batches: whole waves of the densest CTA, then the remainder as its own launch with whatever width is
        // cheapest for it -- e.g. 2,048 ciphertexts = 2 waves of 888 + 272 ciphertexts at KCT = 2 (4.6 ms) instead of a third
        // 8.6 ms wave.  Ciphertext indices stay global (BrArgs.ct_base), so no pointer is offset.
        const unsigned wave = sm_total * dense;
        const unsigned full = (a.B / wave) * wave, tail = a.B - full;
        if (full > 0 && tail > 0) {
            double best_tail = 1e30;
            for (int w = 0; w < n_widths; w++) best_tail = std::min(best_tail, ((tail + sm_total * widths[w] - 1) / (sm_total * widths[w])) * t_cta[w]);
            if (tune.latency_mode != 0 && tail <= sm_total && !a.wide_round) best_tail = std::min(best_tail, 2.5);
            if (best_tail < t_dense - 1e-9) {
                BrArgs m = a, r = a;
                m.B = full;
                r.B = tail;
                r.ct_base = full;
                BrTuning tm = tune;
                tm.kct = dense;
                cudaError_t e = launch_blind_rotate(m, tm, track_margin, s, launches);
                if (e != cudaSuccess) return e;
                return launch_blind_rotate(r, tune, track_margin, s, launches);
            }
        }
    }
    if (kct <= 0) {   // minimise (number of CTA waves) x (time of one CTA at that width)
        double best = 1e30;
        for (int w = 0; w < n_widths; w++) {
            const unsigned waves = (a.B + sm_total * widths[w] - 1) / (sm_total * widths[w]);
            const double cost = waves * t_cta[w];
            if (cost < best - 1e-9) { best = cost; kct = widths[w]; }
        }
    }
    if (launches) (*launches)++;
    // tensor-memory twiddles: the KCT = 6 default, or forced (tuning key "twt" = 1) at 4 and 5
    // (teams of two on top of it -- launch_twt<6, 2>, key loads of two ciphertexts merged -- measured 102.7 k against 103.3 k/s:
    // not instantiated)
    if (twt_ok && (kct == 6 || (tune.twt > 0 && kct >= 4)))
        return kct == 4 ? launch_twt<4>(a, s) : kct == 5 ? launch_twt<5>(a, s) : launch_twt<6>(a, s);
    if (tune.use_tma != 0 && (kct == 2 || kct == 4 || kct == 6) && tune.team == 2)
        return kct == 2 ? launch_team2<2>(a, track_margin, s) : kct == 4 ? launch_team2<4>(a, track_margin, s) : launch_team2<6>(a, track_margin, s);
    switch (kct) {
        case 1: return launch_kct<1>(a, tune.use_tma != 0, track_margin, s);
        case 2: return launch_kct<2>(a, tune.use_tma != 0, track_margin, s);
        case 3: return launch_kct<3>(a, tune.use_tma != 0, track_margin, s);
        case 4: return launch_kct<4>(a, tune.use_tma != 0, track_margin, s);
        case 5: return launch_kct<5>(a, tune.use_tma != 0, track_margin, s);
        default: return launch_kct<6>(a, tune.use_tma != 0, track_margin, s);
    }
}

}  // namespace tfhe_b200

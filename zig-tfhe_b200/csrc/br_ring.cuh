// br_ring.cuh -- the bootstrapping-key ring (cp.async.bulk + mbarriers) and the exchange-buffer state shared by the fast
// (blind_rotate.cu) and the exact (blind_rotate_exact.cu) throughput kernels.
#pragma once
#include "br_common.cuh"
#include "negacyclic_fft.cuh"

namespace tfhe_b200 {

constexpr int kMaxStages = 4;  // deepest key ring (16 KiB per stage)

__host__ __device__ constexpr int align16(int x) { return (x + 15) & ~15; }

// Key-ring producer state, live only in thread 0 of the CTA (see header comment).
struct Producer {
    const cplx *src;      // next chunk in global memory
    cplx *ring;
    uint64_t *full_bar, *empty_bar;
    int remaining;        // chunks still to issue
    int issued;           // chunks issued so far (the first `stages` need no empty wait)
    int stages;
    uint64_t policy;      // L2 evict_last for the key stream
    int stage;
    uint32_t phase;
    bool active;
    // last-arriver refill (ring_release)
    uint32_t *ring_cnt;   // [stages] releases per stage, monotone
    const cplx *bsk;      // chunk 0 of the key
    uint32_t total;       // chunks of the whole blind rotation
    uint32_t ring_warps;  // consumer warps of this CTA
};
__device__ __forceinline__ void producer_poll(Producer &pr) {
    if (pr.active && pr.remaining > 0) {
        if (pr.issued < pr.stages || mbar_test_wait(&pr.empty_bar[pr.stage], pr.phase ^ 1)) {
            mbar_arrive_expect_tx(&pr.full_bar[pr.stage], kBskChunkBytes);
            bulk_g2s(pr.ring + pr.stage * kBskChunkCplx, pr.src, kBskChunkBytes, &pr.full_bar[pr.stage], pr.policy);
            pr.src += kBskChunkCplx;
            pr.remaining--;
            pr.issued++;
            if (++pr.stage == pr.stages) { pr.stage = 0; pr.phase ^= 1; }
        }
    }
}

// Last-arriver refill.  A consumer warp that is done with chunk c (ring stage c mod stages) counts its release with a
// shared-memory atomic; the warp that completes the round -- every consumer warp of the CTA has released the stage --
// issues the bulk copy of chunk c + stages into it at once.  No polling thread, no empty barriers.
// (Measured, profiles/r02_k1_ring.log: a split-phase version that fired the atomic and looked at its return value only at
// the exchange point of the warp's next transform was slower, 92.8 k against 96.2 k bootstraps/s at four ciphertexts per
// CTA -- the refill leaves a few hundred cycles later and two more registers spill.)
__device__ __forceinline__ void ring_release(Producer &pr, int stage, uint32_t c, int lane) {
    __syncwarp();                                  // every lane's reads of the stage are done
    if (lane != 0) return;
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&pr.ring_cnt[stage])) : "memory");
    const uint32_t next = c + (uint32_t)pr.stages;
    if (old + 1u == pr.ring_warps * (next / (uint32_t)pr.stages) && next < pr.total) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy reads before the async-proxy refill
        mbar_arrive_expect_tx(&pr.full_bar[stage], kBskChunkBytes);
        bulk_g2s(pr.ring + stage * kBskChunkCplx, pr.bsk + (size_t)next * kBskChunkCplx, kBskChunkBytes, &pr.full_bar[stage], l2_policy_evict_last());
    }
}

// (Key chunks through TENSOR memory instead of shared memory for the pointwise MAC -- bulk copy into a shared-memory stage,
// sixteen tcgen05.cp.64x128b.warpx2::02_13 per chunk into 64 tensor-memory columns, tcgen05.ld.x16 by every thread from its
// own lane, a second three-stage ring with its own last-arriver protocol -- was built and measured at six ciphertexts per
// CTA: bit-exact, 22 % fewer shared-memory wavefronts, and 76.2 k instead of 103.3 k bootstraps/s.  tcgen05.ld delivers
// ~64-100 B/clk/SM; with the twiddles already coming from tensor memory the key reads (576 KB per SM and step on top of
// 786 KB) saturate it.  Removed again; profiles/r02_k1_ring.log.)

// Per-group exchange state.  X2 is double-buffered when DBX2 (one named barrier per transform instead
// of two): a writer of buffer b at transform k has passed the barrier of transform k-1, which every
// reader of b at transform k-2 reached only after finishing its reads.
struct Xbuf {
    cplx *x1;
    cplx *x2;     // two consecutive buffers of kX2Slots when double-buffered
    int flip;     // 0 or kX2Slots
};

}  // namespace tfhe_b200

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

// Per-group exchange state.  X2 is double-buffered when DBX2 (one named barrier per transform instead
// of two): a writer of buffer b at transform k has passed the barrier of transform k-1, which every
// reader of b at transform k-2 reached only after finishing its reads.
struct Xbuf {
    cplx *x1;
    cplx *x2;     // two consecutive buffers of kX2Slots when double-buffered
    int flip;     // 0 or kX2Slots
};

}  // namespace tfhe_b200

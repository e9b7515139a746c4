// keyswitch_tc.cu -- K2t: identity key switching as an exact integer contraction on the 5th-generation tensor cores.
//
// Replaces trgsw.identityKeySwitching (src/trgsw.zig:471-502) for large batches on the BASEBIT = 2, 4 and 5 sets:
//   res = (0, ..., 0, src.b) - sum_{i < N, j < t} KSK[i][j][k_ij],   k_ij = digit_j(src.a[i] + prec_offset), k_ij != 0
// With the digits expanded to one-hot bytes the sum is a dense product
//   S[ct][col] = sum_K onehot[ct][K] * KSK[K][col]   (K = 3 (i t + j) + k - 1, 27,648 values at the 128-bit set)
// over u32 values mod 2^32.  The key is split into its four byte planes; each plane is an unsigned 8-bit operand of
// tcgen05.mma kind::i8 with 32-bit integer accumulation in tensor memory.  At most N t = 9,216 rows are selected per
// ciphertext, so a plane sum stays below 9,216 x 255 < 2^22: no overflow, and
//   S = P0 + (P1 << 8) + (P2 << 16) + (P3 << 24)  (mod 2^32)
// is the reference's wrapping sum bit for bit (integer arithmetic, any order).  The scalar-pipe version
// (keyswitch.cu) spends 24.6 M thread instructions per ciphertext on a decode -> compare -> branch chain
// (profiles/r01_ncu_k2_splits.txt); here the selection IS the matrix product.
//
// Work decomposition.  One CTA = 128 ciphertexts (MMA M) x one column group of up to 128 output columns x 4 planes
// (up to 512 accumulator columns = all of tensor memory), looping over all K in blocks of 96 bytes (32 (i, j) pairs =
// three K = 32 MMA steps).  Warps 0-3: each thread owns one ciphertext row, expands one 64-bit word of its digit
// stream into 96 one-hot bytes per block and stores them in the canonical K-major no-swizzle operand layout; after the
// last block the same warps read the accumulators back (tcgen05.ld), recombine the planes and store the row.
// Warp 4: one thread streams the pre-arranged key tile of each block (48 KiB, already in operand layout in global
// memory, so a plain cp.async.bulk suffices) into a 3-deep ring.  Warp 5: one thread issues the MMAs and commits
// each stage back to the producers (tcgen05.commit -> mbarrier).
//
// Other bases: BASEBIT = 2 packs a pair into 3 one-hot bytes (k = 1, 2, 3; 32 pairs per 96-byte block).  BASEBIT = 4 / 5 (UINT2 /
// UINT4, t = 3) use 2^BASEBIT bytes per pair, byte k hot for digit k, with an all-zero key row behind byte 0: 6 resp. 3 pairs
// per block, so the block geometry (96 bytes, three K = 32 MMA steps) and everything downstream of the expansion are unchanged.
//
// Operand layout (both operands K-major, SWIZZLE_NONE): 8 rows x 16 bytes core matrices, 128 contiguous bytes each;
// the six core matrices of a row group along K are contiguous (LBO = 128), row groups follow every 768 bytes (SBO).
#include <cuda_runtime.h>

#include "br_common.cuh"
#include "kernels.cuh"

namespace tfhe_b200 {

namespace {

constexpr int kTcRows = 128;                       // ciphertexts per CTA (MMA M)
constexpr int kTcBlockK = 96;                      // one-hot bytes per K block = three K = 32 MMA steps = one 64-bit digit word per ciphertext
__host__ __device__ constexpr int tc_pair_bytes(int basebit) { return basebit == 2 ? 3 : (1 << basebit); }
__host__ __device__ constexpr int tc_pairs_per_block(int basebit) { return kTcBlockK / tc_pair_bytes(basebit); }   // 32, 6 (BASEBIT 4), 3 (BASEBIT 5)
constexpr int kTcCores = kTcBlockK / 16;           // 6 core matrices along K
constexpr int kTcLbo = 128, kTcSbo = kTcCores * 128;
constexpr int kTcGroupCols = 128;                  // output columns per CTA (x 4 planes = 512 accumulator columns)
constexpr int kTcStages = 3;
constexpr int kTcABytes = kTcRows * kTcBlockK;     // 12,288
constexpr int kTcBBytesMax = 4 * kTcGroupCols * kTcBlockK;   // 49,152
constexpr int kTcThreads = 192;

__host__ __device__ constexpr int tc_group_width(int pitch, int g) { return (pitch - g * kTcGroupCols) < kTcGroupCols ? (pitch - g * kTcGroupCols) : kTcGroupCols; }
__host__ __device__ constexpr size_t tc_group_offset(int pitch, int g, int nblocks) {   // byte offset of column group g in the tensor-core key
    return (size_t)g * kTcGroupCols * 4 * kTcBlockK * nblocks;
}
// byte offset inside one operand tile: row r (ciphertext or key column), byte kb of the K block
__host__ __device__ constexpr int tc_tile_offset(int r, int kb) { return (r >> 3) * kTcSbo + (kb >> 4) * kTcLbo + (r & 7) * 16 + (kb & 15); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(kTcLbo >> 4) << 16) | ((uint64_t)(kTcSbo >> 4) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem] * B[smem], unsigned 8-bit x unsigned 8-bit -> s32
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

// digit stream: word w of ciphertext ct holds the BASEBIT-bit digits of the pairs of K block w (pair p = i t + j), pair q of the block at
// bits BASEBIT * q
__global__ void ks_digits_kernel(const uint32_t *__restrict__ lv1, uint64_t *__restrict__ ds, uint32_t B, int t, int words, int basebit) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * words) return;
    const size_t ct = idx / words;
    const int w = (int)(idx - ct * words);
    const int ppb = tc_pairs_per_block(basebit);
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));    // trgsw.zig:483
    const uint32_t kmask = (1u << basebit) - 1u;
    const uint32_t *a = lv1 + ct * (size_t)(kN + 1);
    uint64_t word = 0;
    int p = w * ppb;
    int i = p / t, j = p - i * t;
    uint32_t abar = a[i] + prec_offset;
    for (int q = 0; q < ppb; q++) {
        const uint64_t k = (abar >> (32 - basebit * (j + 1))) & kmask;   // trgsw.zig:488-489
        word |= k << (basebit * q);
        if (++j == t) {
            j = 0;
            i++;
            if (q + 1 < ppb) abar = a[i] + prec_offset;
        }
    }
    ds[idx] = word;
}

// one-time re-layout: packed key [pairs][base - 1][pitch] u32 -> tensor-core key [group][block][operand tile], byte planes split.
// Tile row = plane * W + column-in-group; tile K byte = (pair in block) * pair_bytes + slot, slot = k - 1 (BASEBIT 2) or k (else: slot 0 = zeros).
__global__ void ksk_to_tc_kernel(const uint32_t *__restrict__ ksk, uint8_t *__restrict__ out, int pitch, int nblocks, int basebit) {
    const int g = blockIdx.y, b = blockIdx.x;
    const int W = tc_group_width(pitch, g);
    const int pb = tc_pair_bytes(basebit), ppb = tc_pairs_per_block(basebit), rows_per_pair = (1 << basebit) - 1;
    uint8_t *tile = out + tc_group_offset(pitch, g, nblocks) + (size_t)b * (4 * W * kTcBlockK);
    const int total = 4 * W * kTcBlockK;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int kb = e / (4 * W), r = e - kb * (4 * W);      // key column fastest: coalesced reads of a key row
        const int plane = r / W, cc = r - plane * W;
        const int pair = b * ppb + kb / pb, slot = kb % pb;
        const int krow = basebit == 2 ? slot : slot - 1;       // row of the packed key (digit k = krow + 1), -1: the zero row of digit 0
        const uint32_t v = krow < 0 ? 0u : ksk[((size_t)pair * rows_per_pair + krow) * pitch + g * kTcGroupCols + cc];
        tile[tc_tile_offset(r, kb)] = (uint8_t)(v >> (8 * plane));
    }
}

template <int BASEBIT>
__global__ void __launch_bounds__(kTcThreads, 1)
    keyswitch_tc_kernel(const uint64_t *__restrict__ ds, const uint8_t *__restrict__ ksk_tc, const uint32_t *__restrict__ lv1, uint32_t *__restrict__ lv0,
                        uint32_t B, int n, int pitch, int nblocks) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *a_ring = smem_raw;                                   // [stages][12,288]
    unsigned char *b_ring = smem_raw + kTcStages * kTcABytes;           // [stages][49,152]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(b_ring + kTcStages * kTcBBytesMax);
    uint64_t *empty_bar = full_bar + kTcStages;
    uint64_t *accum_bar = empty_bar + kTcStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x;
    const size_t ct0 = (size_t)blockIdx.y * kTcRows;
    const int W = tc_group_width(pitch, g);
    const int ncols = 4 * W;                       // accumulator columns of this CTA
    const uint32_t b_bytes = (uint32_t)ncols * kTcBlockK;

    if (tid == 0) {
        for (int s = 0; s < kTcStages; s++) {
            mbar_init(&full_bar[s], 1 + 4);        // key tile (expect_tx arrive) + the four expanding warps
            mbar_init(&empty_bar[s], 1);           // tcgen05.commit
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, 512);
        tmem_fence_before_sync();
    }
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ---- one-hot expansion of this thread's ciphertext row, then the epilogue for the same row
        const int r = tid;
        const size_t ct = ct0 + r;
        const bool live = ct < B;
        const uint64_t *my_ds = ds + ct * (size_t)nblocks;
        uint64_t next = live ? __ldg(my_ds) : 0ull;
        unsigned char *row_base = a_ring + (r >> 3) * kTcSbo + (r & 7) * 16;
        for (int kb = 0; kb < nblocks; kb++) {
            const int s = kb % kTcStages;
            const uint32_t ph = (uint32_t)(kb / kTcStages) & 1u;
            const uint64_t dw = next;
            if (live && kb + 1 < nblocks) next = __ldg(my_ds + kb + 1);
            uint32_t w[24];
            if (BASEBIT == 2) {
#pragma unroll
                for (int q4 = 0; q4 < 8; q4++) {       // four pairs -> three words
                    uint32_t v[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const uint32_t k8 = (uint32_t)((dw >> (2 * (4 * q4 + e))) & 3ull) << 3;
                        v[e] = (1u << k8) >> 8;         // k = 0: no byte; k = 1, 2, 3: byte k - 1 of the pair's three
                    }
                    w[3 * q4] = v[0] | (v[1] << 24);
                    w[3 * q4 + 1] = (v[1] >> 8) | (v[2] << 16);
                    w[3 * q4 + 2] = (v[2] >> 16) | (v[3] << 8);
                }
            } else {                                   // 2^BASEBIT bytes per pair, byte k hot (byte 0 meets the zero row)
                constexpr int kWordsPerPair = (1 << BASEBIT) / 4, kPairs = 24 / kWordsPerPair;
#pragma unroll
                for (int p = 0; p < kPairs; p++) {
                    const uint32_t k = (uint32_t)(dw >> (BASEBIT * p)) & ((1u << BASEBIT) - 1u);
                    const uint32_t hot = 1u << (8u * (k & 3u));
#pragma unroll
                    for (int q = 0; q < kWordsPerPair; q++) w[p * kWordsPerPair + q] = ((k >> 2) == (uint32_t)q) ? hot : 0u;
                }
            }
            mbar_wait(&empty_bar[s], ph ^ 1u);
            unsigned char *dst = row_base + s * kTcABytes;
#pragma unroll
            for (int c = 0; c < kTcCores; c++)
                *reinterpret_cast<uint4 *>(dst + c * kTcLbo) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        }
        // ---- epilogue: this row's accumulators (lane = row, 32-bit columns: plane * W + column)
        mbar_wait(accum_bar, 0);
        tmem_fence_after_sync();
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t *o = lv0 + ct * (size_t)(n + 1);
        const uint32_t body = live ? lv1[ct * (size_t)(kN + 1) + kN] : 0u;
        for (int c0 = 0; c0 < W; c0 += 8) {
            uint32_t p0[8], p1[8], p2[8], p3[8];
            tmem_ld8(trow + (uint32_t)c0, p0);
            tmem_ld8(trow + (uint32_t)(W + c0), p1);
            tmem_ld8(trow + (uint32_t)(2 * W + c0), p2);
            tmem_ld8(trow + (uint32_t)(3 * W + c0), p3);
            tmem_wait_ld();
            if (live) {
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int col = g * kTcGroupCols + c0 + e;
                    if (col > n) continue;
                    const uint32_t sum = p0[e] + (p1[e] << 8) + (p2[e] << 16) + (p3[e] << 24);
                    o[col] = (col == n ? body : 0u) - sum;      // trgsw.zig:481, 494-497
                }
            }
        }
        tmem_fence_before_sync();
    } else if (warp == 4) {
        // ---- key tiles: one bulk copy per K block
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_last();
            const uint8_t *src = ksk_tc + tc_group_offset(pitch, g, nblocks);
            for (int kb = 0; kb < nblocks; kb++) {
                const int s = kb % kTcStages;
                const uint32_t ph = (uint32_t)(kb / kTcStages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], b_bytes);
                bulk_g2s(b_ring + s * kTcBBytesMax, src + (size_t)kb * b_bytes, b_bytes, &full_bar[s], policy);
            }
        }
        __syncwarp();
    } else {
        // ---- MMA issue: per K block three K = 32 steps x up to two N <= 256 halves
        if (lane == 0) {
            const int n0 = ncols > 256 ? 256 : ncols, n1 = ncols - n0;
            const uint32_t idesc_hi = (2u << 4) | ((uint32_t)(kTcRows >> 4) << 24);    // s32 accumulate, u8 x u8, both K-major, M = 128
            const uint32_t idesc0 = idesc_hi | ((uint32_t)(n0 >> 3) << 17);
            const uint32_t idesc1 = idesc_hi | ((uint32_t)(n1 >> 3) << 17);
            for (int kb = 0; kb < nblocks; kb++) {
                const int s = kb % kTcStages;
                const uint32_t ph = (uint32_t)(kb / kTcStages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tmem_fence_after_sync();
                const uint32_t a_addr = smem_u32(a_ring + s * kTcABytes), b_addr = smem_u32(b_ring + s * kTcBBytesMax);
#pragma unroll
                for (int ks = 0; ks < 3; ks++) {
                    const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
                    const uint64_t da = smem_desc(a_addr + ks * 2 * kTcLbo);
                    mma_i8(tmem_base, da, smem_desc(b_addr + ks * 2 * kTcLbo), idesc0, acc);
                    if (n1 > 0) mma_i8(tmem_base + 256u, da, smem_desc(b_addr + (256 / 8) * kTcSbo + ks * 2 * kTcLbo), idesc1, acc);
                }
                mma_commit(&empty_bar[s]);          // the stage is free once these MMAs have read it
            }
            mma_commit(accum_bar);                  // all accumulators complete
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 5) {
        tmem_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

bool keyswitch_tc_supported(int basebit, int iks_t, int in_dim, int pitch) {
    if (basebit != 2 && basebit != 4 && basebit != 5) return false;      // pair_bytes must divide the 96-byte block
    return in_dim == kN && (kN * iks_t) % tc_pairs_per_block(basebit) == 0 && basebit * tc_pairs_per_block(basebit) <= 64 && pitch % 4 == 0 && pitch >= 16;
}
size_t keyswitch_tc_digit_words(int basebit, int iks_t) { return keyswitch_tc_supported(basebit, iks_t, kN, 16) ? (size_t)kN * iks_t / tc_pairs_per_block(basebit) : 0; }
size_t keyswitch_tc_key_bytes(int pitch, int basebit, int iks_t) { return keyswitch_tc_digit_words(basebit, iks_t) * kTcBlockK * (size_t)pitch * 4; }

cudaError_t launch_ksk_to_tc(const uint32_t *ksk_packed, uint8_t *out, int basebit, int iks_t, int pitch, cudaStream_t s, uint64_t *launches) {
    const int nblocks = (int)keyswitch_tc_digit_words(basebit, iks_t);
    const int groups = (pitch + kTcGroupCols - 1) / kTcGroupCols;
    ksk_to_tc_kernel<<<dim3(nblocks, groups), 256, 0, s>>>(ksk_packed, out, pitch, nblocks, basebit);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_keyswitch_tc(const KsArgs &a, const uint8_t *ksk_tc, uint64_t *digits, cudaStream_t s, uint64_t *launches) {
    if (a.B == 0) return cudaSuccess;
    const int nblocks = (int)keyswitch_tc_digit_words(a.basebit, a.iks_t);
    const int groups = (a.pitch + kTcGroupCols - 1) / kTcGroupCols;
    const size_t words = (size_t)a.B * nblocks;
    ks_digits_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(a.lv1, digits, a.B, a.iks_t, nblocks, a.basebit);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)kTcStages * (kTcABytes + kTcBBytesMax) + (2 * kTcStages + 1) * 8 + 16;
    auto kern = a.basebit == 2 ? keyswitch_tc_kernel<2> : a.basebit == 4 ? keyswitch_tc_kernel<4> : keyswitch_tc_kernel<5>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const dim3 grid(groups, (a.B + kTcRows - 1) / kTcRows);
    kern<<<grid, kTcThreads, smem, s>>>(digits, ksk_tc, a.lv1, a.lv0, a.B, a.n, a.pitch, nblocks);
    if (launches) (*launches) += 2;
    return cudaGetLastError();
}

}  // namespace tfhe_b200

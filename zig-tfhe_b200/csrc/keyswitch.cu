// keyswitch.cu -- K2: batched identity key switching lv1 -> lv0 (integer, bit-exact on any schedule); the same kernel
// with in_dim = n is the batched proxy re-encryption lv0 -> lv0 (proxy_reenc.reencryptTLWELv0, src/proxy_reenc.zig:267-306).
//
// Replaces trgsw.identityKeySwitching (src/trgsw.zig:471-502):
//   res.b = src.b;  for i < N, j < t:  k = digit_j(src.a[i] + prec_offset);  if k != 0: res -= KSK[i][j][k]
// u32 wrapping subtraction is associative and commutative, so any tiling / split / atomic order gives
// the reference's bits.
//
// Tiling: one CTA owns CT ciphertexts x all n+1 output columns (each thread 4 columns, CT uint4
// accumulators in registers) and walks a slice of the (i, j) pairs.  For every pair the base-1 key rows
// are loaded once (coalesced 16-byte loads) and applied to all CT ciphertexts, so key traffic is
// 1/CT of the per-ciphertext gather (19.4 MB/ciphertext at the 128-bit set).  Small batches split the
// i range over several CTAs and combine with u32 atomics so the grid still fills the 148 SMs.
//
// Device key layout (built by repack_ksk_kernel): [N][t][base-1][pitch] u32, pitch = roundup4(n+1);
// the never-read k = 0 rows of the reference layout (src/key.zig:158-161) are dropped.
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace tfhe_b200 {

namespace {

__device__ __forceinline__ void sub4(uint4 &a, const uint4 &r) {
    a.x -= r.x; a.y -= r.y; a.z -= r.z; a.w -= r.w;
}

// BASEBIT = 2 fast path (80/110/128-bit and UINT1 sets): 3 rows per (i,j) held in registers.
// BASEBIT = 0: generic base, row fetched per ciphertext.
// V = uint4 vectors per thread (thread `col` owns columns 4*col.. and, for V = 2, 4*(col + pitch4/2)..): the digit
// decode + warp-uniform branch of every (ciphertext, i, j) is amortised over 4*V subtractions.
template <int CT, int BASEBIT, int V>
__global__ void __launch_bounds__(320) keyswitch_kernel(const KsArgs P, int i_per_split, int use_atomics) {
    extern __shared__ uint32_t abar[];   // [CT][i_per_split]
    const int n = P.n, t = P.iks_t;
    const int basebit = BASEBIT ? BASEBIT : P.basebit;
    const int rows_per_pair = (1 << basebit) - 1;
    const int pitch4 = P.pitch >> 2;
    const int half4 = (pitch4 + V - 1) / V;            // uint4 columns per vector slot
    const int col = threadIdx.x;                       // first uint4 column of this thread
    const size_t ct0 = (size_t)blockIdx.x * CT;
    const int i0 = blockIdx.y * i_per_split;
    const int in_dim = P.in_dim;
    const int i1 = min(in_dim, i0 + i_per_split);
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));   // trgsw.zig:483
    const uint32_t kmask = (1u << basebit) - 1u;

    for (int idx = threadIdx.x; idx < CT * (i1 - i0); idx += blockDim.x) {
        const int c = idx / (i1 - i0), i = idx - c * (i1 - i0);
        const size_t ct = ct0 + c;
        // inactive slots: abar = 0 -> every digit is 0 -> nothing subtracted
        abar[c * i_per_split + i] = (ct < P.B) ? P.lv1[ct * (size_t)(in_dim + 1) + i0 + i] + prec_offset : 0u;
    }
    __syncthreads();
    if (col >= half4) return;
    bool live[V];
#pragma unroll
    for (int v = 0; v < V; v++) live[v] = col + v * half4 < pitch4;

    uint4 acc[CT][V];
#pragma unroll
    for (int c = 0; c < CT; c++)
#pragma unroll
        for (int v = 0; v < V; v++) acc[c][v] = make_uint4(0u, 0u, 0u, 0u);

    const uint4 *ksk4 = reinterpret_cast<const uint4 *>(P.ksk);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    // every CTA walks its i range from a different starting point (the subtraction order is free): CTAs that start
    // together do not pull the same key rows through the same L2 slices at the same time
    const int span = i1 - i0;
    int i = i0 + (P.rot >= 0 && span > 0 ? (int)((blockIdx.x * 37u) % (unsigned)span) : 0);
    for (int k = 0; k < span; k++, i = (i + 1 == i1) ? i0 : i + 1) {
        uint32_t ab[CT];
#pragma unroll
        for (int c = 0; c < CT; c++) ab[c] = abar[c * i_per_split + (i - i0)];
        const uint4 *rowp = ksk4 + ((size_t)i * t) * rows_per_pair * pitch4 + col;
        for (int j = 0; j < t; j++, rowp += (size_t)rows_per_pair * pitch4) {
            const int sh = 32 - (j + 1) * basebit;
            if (BASEBIT == 2) {
                uint4 r1[V], r2[V], r3[V];
#pragma unroll
                for (int v = 0; v < V; v++) {
                    r1[v] = live[v] ? __ldg(rowp + v * half4) : zero4;
                    r2[v] = live[v] ? __ldg(rowp + v * half4 + pitch4) : zero4;
                    r3[v] = live[v] ? __ldg(rowp + v * half4 + 2 * pitch4) : zero4;
                }
#pragma unroll
                for (int c = 0; c < CT; c++) {
                    const uint32_t k = (ab[c] >> sh) & 3u;     // warp-uniform
                    if (k == 1u) {
#pragma unroll
                        for (int v = 0; v < V; v++) sub4(acc[c][v], r1[v]);
                    } else if (k == 2u) {
#pragma unroll
                        for (int v = 0; v < V; v++) sub4(acc[c][v], r2[v]);
                    } else if (k == 3u) {
#pragma unroll
                        for (int v = 0; v < V; v++) sub4(acc[c][v], r3[v]);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < CT; c++) {
                    const uint32_t k = (ab[c] >> sh) & kmask;
                    if (k != 0u) {
#pragma unroll
                        for (int v = 0; v < V; v++)
                            if (live[v]) sub4(acc[c][v], __ldg(rowp + v * half4 + (size_t)(k - 1u) * pitch4));
                    }
                }
            }
        }
    }

    // write back: column n additionally receives src.b (trgsw.zig:481), added by the first split only
#pragma unroll
    for (int c = 0; c < CT; c++) {
        const size_t ct = ct0 + c;
        if (ct >= P.B) break;
        uint32_t *o = P.lv0 + ct * (size_t)(n + 1);
#pragma unroll
        for (int v = 0; v < V; v++) {
            const uint32_t vals[4] = {acc[c][v].x, acc[c][v].y, acc[c][v].z, acc[c][v].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int x = (col + v * half4) * 4 + e;
                if (!live[v] || x > n) continue;
                uint32_t val = vals[e];
                if (x == n && blockIdx.y == 0) val += P.lv1[ct * (size_t)(in_dim + 1) + in_dim];
                if (use_atomics) atomicAdd(&o[x], val);
                else o[x] = val;
            }
        }
    }
}

template <int CT, int V>
cudaError_t launch_ct(const KsArgs &a, int splits, cudaStream_t s) {
    const int i_per_split = (a.in_dim + splits - 1) / splits;
    const size_t smem = (size_t)CT * i_per_split * sizeof(uint32_t);
    const dim3 grid((a.B + CT - 1) / CT, splits);
    const int pitch4 = a.pitch >> 2;
    const int threads = ((((pitch4 + V - 1) / V) + 31) / 32) * 32;
    cudaError_t e;
    if (a.basebit == 2) {
        auto k = keyswitch_kernel<CT, 2, V>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, threads, smem, s>>>(a, i_per_split, splits > 1);
    } else {
        auto k = keyswitch_kernel<CT, 0, V>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k<<<grid, threads, smem, s>>>(a, i_per_split, splits > 1);
    }
    return cudaGetLastError();
}

__global__ void repack_ksk_kernel(const uint32_t *ref, size_t ref_stride, uint32_t *out, int n, int base, int pairs, int pitch) {
    // one block per (i,j) pair, rows k = 1..base-1
    const int pair = blockIdx.x;
    if (pair >= pairs) return;
    for (int k = 1; k < base; k++) {
        const uint32_t *src = ref + ((size_t)pair * base + k) * ref_stride;
        uint32_t *dst = out + ((size_t)pair * (base - 1) + (k - 1)) * pitch;
        for (int x = threadIdx.x; x < pitch; x += blockDim.x) dst[x] = (x <= n) ? src[x] : 0u;
    }
}

__global__ void negate_kernel(const uint32_t *a, uint32_t *out, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = 0u - a[i];
}

// trivial ciphertexts of Gates.constant (src/gates.zig:144-151): rows of w words, mask 0, body `body`
__global__ void fill_constant_kernel(uint32_t *rows, size_t count, int w, uint32_t body) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) rows[i] = ((int)(i % (size_t)w) == w - 1) ? body : 0u;
}

// trlwe.sampleExtractIndex(., k) (src/trlwe.zig:146-162): p[i] = a[k - i] for i <= k, -a[N + k - i] for i > k; p[N] = b[k]
__global__ void sample_extract_kernel(const uint32_t *trlwe, uint32_t *lv1, uint32_t B, int k) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * (kN + 1)) return;
    const size_t ct = idx / (kN + 1);
    const int i = (int)(idx - ct * (kN + 1));
    const uint32_t *a = trlwe + ct * (size_t)(2 * kN), *b = a + kN;
    lv1[idx] = (i == kN) ? b[k] : (i <= k) ? a[k - i] : 0u - a[kN + k - i];
}

__global__ void extract2_kernel(const uint32_t *lv1, uint32_t *out, uint32_t B, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t w = (size_t)n + 1;
    if (idx >= (size_t)B * w) return;
    const size_t ct = idx / w;
    const int x = (int)(idx - ct * w);
    // sampleExtractIndex2(., 0): first n mask entries + body (trlwe.zig:165-180)
    out[idx] = (x == n) ? lv1[ct * (size_t)(kN + 1) + kN] : lv1[ct * (size_t)(kN + 1) + x];
}

}  // namespace

cudaError_t launch_keyswitch(const KsArgs &a, int sm_count, cudaStream_t s, uint64_t *launches) {
    if (a.B == 0) return cudaSuccess;
    if (a.pitch > 320 * 4) return cudaErrorInvalidValue;
    // Tile and split, measured on B200 (tools/ks_bench.py, profiles/r01_ks_split_sweep.log):
    //  * tile 8 from 32 ciphertexts up (at B = 65,536: tile 8 = 60 ms, 4 = 70 ms, 16 = 83 ms; at B = 74..2,048 tile 4 re-reads the
    //    key through L2 once per tile and is 1.5-2x slower than tile 8);
    //  * the i range is split over blockIdx.y until the grid holds `fill` CTAs per SM, partial results combined with u32
    //    atomics (order-free): 64 per SM from 1,024 ciphertexts up -- also at B = 65,536, where two splits take 51 ms against 60 ms
    //    for one (shorter CTAs: 19 instead of 9.2 waves of 6 resident CTAs, and different i ranges in flight at once) -- and 16
    //    below that (B = 148: 0.18 ms against 0.67 ms with the 2 per SM of the first version).
    int ct = (a.B >= 32u) ? 8 : 4;
    if (a.tile == 4 || a.tile == 8 || a.tile == 16) ct = a.tile;
    const int tiles = (a.B + ct - 1) / ct;
    int splits = 1;
    // (up to 256 splits: a single ciphertext then walks 4 mask entries per CTA instead of a 288-step dependent chain)
    const int fill = a.fill > 0 ? a.fill : (a.B >= 1024u ? 64 : 16);
    while ((long long)tiles * splits < (long long)fill * sm_count && splits < 256) splits *= 2;
    cudaError_t e;
    if (splits > 1) {
        e = cudaMemsetAsync(a.lv0, 0, (size_t)a.B * (a.n + 1) * sizeof(uint32_t), s);
        if (e != cudaSuccess) return e;
    }
    if (launches) (*launches)++;
    // measured at B = 65,536 (tools/ks_bench.py): tile 8 x 1 vector 60 ms, tile 8 x 2 vectors 79 ms -> one uint4 per thread
    const bool wide = a.vec == 2;
    switch (ct) {
        case 16: return launch_ct<16, 1>(a, splits, s);
        case 8: return wide ? launch_ct<8, 2>(a, splits, s) : launch_ct<8, 1>(a, splits, s);
        default: return wide ? launch_ct<4, 2>(a, splits, s) : launch_ct<4, 1>(a, splits, s);
    }
}

cudaError_t launch_repack_ksk(const uint32_t *ref_ksk, size_t ref_row_stride_u32, uint32_t *out, int n, int basebit, int iks_t,
                              int pitch, int in_dim, cudaStream_t s, uint64_t *launches) {
    const int pairs = in_dim * iks_t;
    repack_ksk_kernel<<<pairs, 256, 0, s>>>(ref_ksk, ref_row_stride_u32, out, n, 1 << basebit, pairs, pitch);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_negate(const uint32_t *a, uint32_t *out, size_t count, cudaStream_t s, uint64_t *launches) {
    if (!count) return cudaSuccess;
    negate_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(a, out, count);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_fill_constant(uint32_t *rows, size_t n_rows, int w, uint32_t body, cudaStream_t s, uint64_t *launches) {
    if (!n_rows) return cudaSuccess;
    const size_t count = n_rows * (size_t)w;
    fill_constant_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(rows, count, w, body);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_sample_extract(const uint32_t *trlwe, uint32_t *lv1, uint32_t B, int k, cudaStream_t s, uint64_t *launches) {
    if (!B) return cudaSuccess;
    const size_t total = (size_t)B * (kN + 1);
    sample_extract_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(trlwe, lv1, B, k);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

cudaError_t launch_extract2(const uint32_t *lv1, uint32_t *out, uint32_t B, int n, cudaStream_t s, uint64_t *launches) {
    if (!B) return cudaSuccess;
    const size_t total = (size_t)B * (n + 1);
    extract2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(lv1, out, B, n);
    if (launches) (*launches)++;
    return cudaGetLastError();
}

}  // namespace tfhe_b200

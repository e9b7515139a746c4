// br_common.cuh -- device helpers shared by the fast and exact blind-rotation kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfhe_b200 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking phase test (try_wait may suspend the thread for a system-dependent time: not for polling loops that
// have other work to do)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// L2 cache policy: keep these lines (the bootstrapping key is re-read by every CTA wave; batch I/O streams past it)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// ---- Tensor Memory (TMEM) as per-lane private storage --------------------------------------------
// TMEM (256 KiB per SM, 128 lanes x 512 32-bit columns) is reachable only through tcgen05 instructions;
// with the 32x32b shape every thread of a warp reads/writes its own lane, so it is a register-file
// extension with its own data path (it never touches the shared-memory pipe).  K1 keeps the 2 x 8
// complex MAC accumulators there when the register budget is 168 (KCT = 5, 6).
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_slot, uint32_t ncols) {   // whole warp, converged
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // whole warp, converged
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}

// modulus switch 2^32 -> 2N of blindRotate (src/trgsw.zig:297,312): (x + 2^20) >> 21 in 64 bits, so 2N itself can occur.
// shift > 0: the result is a multiple of 2^shift (rounded to it) -- the coarser switch of a many-function bootstrap, whose
// test vector interleaves 2^shift functions and whose accumulator is sample-extracted at indices 0 .. 2^shift - 1.
__device__ __forceinline__ uint32_t mod_switch_2n(uint32_t x, int shift) {
    return (uint32_t)((((unsigned long long)x + (1ull << (20 + shift))) >> (21 + shift)) << shift);
}

// linear part of the ten bootstrapped gates (gates.zig:48-121); op < 0: identity (plain bootstrap)
__device__ __forceinline__ uint32_t gate_linear(int op, uint32_t a, uint32_t b) {
    switch (op) {
        case 0: case 5: return 0u - a - b;   // NAND, NOR      : -a - b
        case 1: case 2: return a + b;        // OR, AND        :  a + b
        case 3: return a + 2u * b;           // XOR            :  a + 2b   (addMul, gates.zig:72)
        case 4: return a - 2u * b;           // XNOR           :  a - 2b   (subMul, gates.zig:79)
        case 6: case 8: return b - a;        // ANDNY, ORNY    : -a + b
        case 7: case 9: return a - b;        // ANDYN, ORYN    :  a - b
        default: return a;
    }
}
// constant added to the body: f64ToTorus(+-1/8, +-1/4) (utils.zig:28-33)
__device__ __forceinline__ uint32_t gate_constant(int op) {
    switch (op) {
        case 0: case 1: case 8: case 9: return 0x20000000u;   // +1/8
        case 2: case 5: case 6: case 7: return 0xE0000000u;   // -1/8
        case 3: return 0x40000000u;                            // +1/4
        case 4: return 0xC0000000u;                            // -1/4
        default: return 0u;
    }
}

// Operands of item `ct`: plain batch (rows ct of in_a / in_b) or one gate of a circuit level (rows of the wire store).
// neg bit 0 / 1: operand a / b enters through Gates.notGate (src/gates.zig:131-133), i.e. negated coefficient-wise.
struct GateOperands {
    const uint32_t *a, *b;
    int op;
    uint32_t neg;
};
template <class Args>
__device__ __forceinline__ GateOperands gate_operands(const Args &P, size_t ct, int n) {
    GateOperands g;
    const size_t w = (size_t)n + 1;
    if (P.lvl_a) {
        const uint32_t gi = (uint32_t)(ct / P.inst), k = (uint32_t)(ct - (size_t)gi * P.inst);
        const uint32_t wa = P.lvl_a[gi], wb = P.lvl_b[gi];
        g.op = P.lvl_ops[gi];
        g.a = P.in_a + ((size_t)(wa & 0x7fffffffu) * P.inst + k) * w;
        g.b = P.in_a + ((size_t)(wb & 0x7fffffffu) * P.inst + k) * w;
        g.neg = (wa >> 31) | ((wb >> 31) << 1);
    } else {
        g.op = P.ops ? P.ops[ct] : P.op;
        g.a = P.in_a + ct * w;
        g.b = (g.op >= 0) ? P.in_b + ct * w : g.a;
        g.neg = 0u;
    }
    return g;
}
__device__ __forceinline__ uint32_t gate_linear_signed(const GateOperands &g, int i) {
    const uint32_t a = g.a[i], b = g.b[i];
    return gate_linear(g.op, (g.neg & 1u) ? 0u - a : a, (g.neg & 2u) ? 0u - b : b);
}

}  // namespace tfhe_b200

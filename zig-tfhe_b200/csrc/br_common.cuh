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
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA engine)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
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


}  // namespace tfhe_b200

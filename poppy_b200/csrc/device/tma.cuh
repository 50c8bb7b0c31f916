// mbarrier / TMA primitives shared by the tiled kernels (inline PTX; sm_100a).
#pragma once

#include "common.cuh"
#include "kernels.cuh"

namespace poppy {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    // try_wait with a suspend-time hint: a warp whose stage has not landed yet is parked by the hardware instead of
    // spinning through the issue slots of the warps that have work
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "POPPY_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra POPPY_MBAR_DONE;\n"
        "bra POPPY_MBAR_WAIT;\n"
        "POPPY_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16; both addresses 16-byte aligned), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const TmaMap* map, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace poppy

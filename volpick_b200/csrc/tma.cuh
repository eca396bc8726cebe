// TMA (cp.async.bulk / cp.async.bulk.tensor) helpers for sm_100a: tensor maps are encoded on the host through the driver
// entry point (no link-time dependency on libcuda), passed to kernels as __grid_constant__ parameters.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace vp {

// Host: tiled tensor map over 16-bit elements.  dims / box: innermost first; strides_bytes[i] = byte stride of dimension
// i + 1 (multiples of 16).  No swizzle, no interleave, out-of-range elements read as zero.
int tma_encode_u16(CUtensorMap *map, const void *gaddr, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                   const uint32_t *box);

#ifdef __CUDACC__
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
// One box of a 5-D tensor map -> shared memory (dst 128-byte aligned); completes `bytes of the box` on the mbarrier.
__device__ __forceinline__ void tma_load_5d(uint32_t dst_smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            dst_smem),
        "l"(reinterpret_cast<uint64_t>(map)), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// Contiguous global -> shared copy (bytes a multiple of 16, both addresses 16-byte aligned).
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif

}  // namespace vp

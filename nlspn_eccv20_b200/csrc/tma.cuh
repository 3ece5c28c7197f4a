// tma.cuh -- minimal TMA (cp.async.bulk.tensor) + mbarrier helpers, inline PTX for sm_100a.
//
// The state plane a propagation iteration gathers from is moved into shared memory as one
// halo'd 2-D box per CTA by the Tensor Memory Accelerator.  TMA's out-of-bounds ZERO FILL is
// exactly the zero padding of the reference's bilinear sampler (modulated_deform_im2col_cuda.cuh
// :37-48), so every per-corner image-border guard disappears from the gather loop.
#pragma once
#include <cuda.h>
// (this header is also included by the non-TMA kernels for the PDL helpers)
#include <cuda_runtime.h>
#include <stdint.h>

namespace nlspn {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// order prior generic-proxy accesses to shared memory before subsequent async-proxy writes
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 3-D tiled load: coordinates (x = fastest/W, y = H, z = plane); signed, may start out of bounds
__device__ __forceinline__ void load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y,
                                        int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// 4-D tiled load: coordinates (x = W, y = H, c = channel, n = batch image)
__device__ __forceinline__ void load_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y,
                                        int c, int n)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c),
        "r"(n)
        : "memory");
}

__device__ __forceinline__ void prefetch_descriptor(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Programmatic dependent launch (PDL): wait for the prerequisite grid / allow dependents
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

} // namespace tma
} // namespace nlspn

// cra_tma.cuh -- bulk asynchronous copies (the TMA unit, SASS UBLKCP) with mbarrier completion.
// A particle image is one contiguous run of nx*nx floats in HBM, so its shared-memory tile is filled by ONE
// cp.async.bulk issued by one thread: no LSU instructions, no register staging, and the copy runs under the
// CTA's table set-up.  Requirements of the instruction: 16-byte aligned source and destination, size a multiple of 16.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cratma {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// one thread: initialise the barrier for `count` arrivals and make it visible to the asynchronous proxy
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// one thread: announce `bytes` and start the global -> shared bulk copy that will complete them on `bar`
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    const unsigned b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}

// any thread: wait until the phase with the given parity has completed (the data is then visible to this thread)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "CRA_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra CRA_MBAR_DONE;\n"
        "bra CRA_MBAR_WAIT;\n"
        "CRA_MBAR_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ bool bulk_ok(const void* gmem_src, size_t bytes)
{
    return ((reinterpret_cast<uintptr_t>(gmem_src) | bytes) & 15) == 0;
}

}  // namespace cratma

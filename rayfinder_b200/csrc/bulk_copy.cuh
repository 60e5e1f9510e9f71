// Bulk asynchronous copies global -> shared memory through the TMA unit (cp.async.bulk, SASS UBLKCP) and the mbarrier
// (SASS SYNCS) that signals their completion — the sm_90+/sm_100a replacement for per-lane loads when a warp wants a
// contiguous block of memory staged in shared memory while it keeps computing.  Used by the warp-per-ray tail kernel
// (straggler.cuh) to fetch the next 1 KB window of BVH nodes while the warp still walks the current one.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace rfb200
{
__device__ __forceinline__ std::uint32_t sharedAddress(const void* p) { return static_cast<std::uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbarrierInit(const std::uint32_t bar, const std::uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
}
// Makes freshly initialised barriers (and earlier generic-proxy accesses to shared memory) visible to the async proxy.
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One arrival that also announces `bytes` of asynchronous traffic the barrier has to wait for.
__device__ __forceinline__ void mbarrierArriveExpectTx(const std::uint32_t bar, const std::uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`.
__device__ __forceinline__ void bulkCopyGlobalToShared(const std::uint32_t dst, const void* src, const std::uint32_t bytes, const std::uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// True once the phase with parity `parity` of `bar` has completed (hardware-suspended wait with a time limit).
__device__ __forceinline__ bool mbarrierTryWait(const std::uint32_t bar, const std::uint32_t parity)
{
    std::uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
// Non-blocking probe of the same condition.
__device__ __forceinline__ bool mbarrierTestWait(const std::uint32_t bar, const std::uint32_t parity)
{
    std::uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
} // namespace rfb200

// The block-local subtrees of the GPU BVH build (k_bvh_build_local, bvh_build_device.cuh) as a translation unit of their own:
// compiled WITHOUT -dlcm=cg, so that the loads of a block hit the L1 of the SM that wrote the data (the block itself).
#include "bvh_build_device.cuh"

namespace rfb200
{
void launchBvhBuildLocal(int grid, cudaStream_t stream, std::uint32_t n, const void* prims, void* order0, void* order1, void* owner, void* slotLeft, void* slotRight,
                         void* counters, void* nodes, void* accum, void* buckets, void* flags, void* scan, void* control, void* leafStart, const void* deferList)
{
    k_bvh_build_local<<<grid, BUILD_THREADS, 0, stream>>>(
        n, static_cast<const Prim*>(prims), static_cast<std::uint32_t*>(order0), static_cast<std::uint32_t*>(order1), static_cast<std::uint32_t*>(owner),
        static_cast<std::uint32_t*>(slotLeft), static_cast<std::uint32_t*>(slotRight), static_cast<std::uint32_t*>(counters), static_cast<BuildNode*>(nodes),
        static_cast<NodeAccum*>(accum), static_cast<BucketAccum*>(buckets), static_cast<unsigned long long*>(flags), static_cast<unsigned long long*>(scan),
        static_cast<FusedControl*>(control), static_cast<std::uint32_t*>(leafStart), static_cast<const std::uint32_t*>(deferList));
}
} // namespace rfb200

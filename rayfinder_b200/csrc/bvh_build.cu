// buildBvh (common/bvh.cpp:81-291) on the GPU, byte-identical to the host builder (host_core.cpp, rf_build_bvh),
// which the CPU test-suite pins byte for byte to the reference's own translation unit.
// SURVEY.md §8(f)-4.
//
// The reference recurses depth-first over a vector it partitions in place.  Here the tree grows LEVEL BY LEVEL
// over one array of primitive positions; every step is data-parallel over the primitives or over the nodes of
// the level:
//
//   boxes      node AABB and centroid AABB of every node of the level: one atomic min/max per primitive and
//              component (warp-aggregated while a warp lies inside one node).  The reference folds the boxes
//              sequentially with `(b < a) ? b : a`, so among numerically equal values (-0.0f / +0.0f) the FIRST one
//              in sequence order survives, and its sign ends up in the .pt file: the node box is reduced on 64-bit
//              keys (ordered value, position) to reproduce that; see loKey() / hiKey().
//   decide     per node: leaf (zero area, identical centroids, one primitive), median split of two primitives
//              (std::nth_element on two elements = one conditional swap), or binned SAH
//   buckets    per primitive of a SAH node: bucket counter and bucket box (atomics)
//   sweep      per SAH node: the 11 candidate costs, literally the host code (bvh_common.h): leaf or split
//   partition  std::partition as libstdc++ implements it for bidirectional iterators: the k-th element from the
//              left that fails the predicate is swapped with the k-th element from the right that satisfies it.
//              The ranks k come from one exclusive scan over (fails, satisfies) flag pairs of all primitives;
//              elements already on their side stay where they are.  Same permutation, no sequential loop.
//   finally    subtree sizes bottom-up, depth-first (pre-order) node numbers top-down — first child = idx + 1,
//              second child = idx + 1 + size(first subtree) — and the 48-byte BvhNode records.
//
// The host only reads back one counter per level (how many nodes the next level has).
#include "bvh_build_device.cuh"

#include <cub/device/device_scan.cuh>

#include <mutex>

namespace rfb200
{
namespace
{
// A device array that keeps its storage between calls and only grows (the builder's workspace, see g_workspace).
template<typename T>
struct Buf
{
    T*          ptr = nullptr;
    std::size_t capacity = 0;
    ~Buf() { release(); }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr, capacity = 0;
    }
    cudaError_t allocate(std::size_t count)
    {
        count = std::max<std::size_t>(count, 1);
        if (count <= capacity) return cudaSuccess;
        release();
        const cudaError_t err = cudaMalloc(&ptr, count * sizeof(T));
        if (err == cudaSuccess) capacity = count;
        return err;
    }
};
inline unsigned gridOf(std::uint64_t items) { return static_cast<unsigned>((items + BUILD_THREADS - 1) / BUILD_THREADS); }
} // namespace
} // namespace rfb200

using namespace rfb200;

namespace
{
// The builder's device arrays, kept between calls (allocating and freeing ~130 MB of them was most of a call: ~25 of 29 ms for
// Sponza); one workspace per process, guarded by a mutex, dropped by rf_build_bvh_device_release or when the device changes.
struct BuildWorkspace
{
    int                     device = -1;
    Buf<rf_positions>       dTris;
    Buf<Prim>               prims;
    Buf<std::uint32_t>      order[2], owner, slotLeft, slotRight, counters, levelStartDev, leafStart, deferList;
    Buf<BuildNode>          nodes;
    Buf<NodeAccum>          accum;
    Buf<BucketAccum>        buckets;
    Buf<unsigned long long> flags, scan, dIndices, blockTotals;
    Buf<rf_bvh_node>        dOut;
    Buf<unsigned char>      scanTemp;
    Buf<FusedControl>       control;
    void release()
    {
        dTris.release(), prims.release(), order[0].release(), order[1].release(), owner.release(), slotLeft.release(), slotRight.release(), counters.release();
        levelStartDev.release(), leafStart.release(), deferList.release(), nodes.release(), accum.release(), buckets.release(), flags.release(), scan.release();
        dIndices.release(), blockTotals.release(), dOut.release(), scanTemp.release(), control.release();
        device = -1;
    }
};
BuildWorkspace g_workspace;
std::mutex     g_workspaceMutex;
float g_lastPhaseMs[12] = {};
std::uint32_t g_lastLevels = 0;
bool g_levelKernels = false; // rf_build_bvh_device_set_mode(1): the level-by-level path (one launch per phase and level), kept for A/B timing
}
extern "C" void rf_build_bvh_device_set_mode(std::int32_t levelKernels) { g_levelKernels = levelKernels != 0; }
// Diagnostics of the last single-launch build: milliseconds block 0 spent in each phase (summed over the levels) and the
// number of levels.  out_phase_ms has 12 entries: boxes, decide, buckets, sweep, scan, offsets, pair, permute, level
// bookkeeping, numbering, emit, unused.
extern "C" void rf_build_bvh_device_release(void)
{
    std::lock_guard<std::mutex> lock(g_workspaceMutex);
    if (g_workspace.device >= 0 && cudaSetDevice(g_workspace.device) == cudaSuccess) g_workspace.release();
    g_workspace.device = -1;
}
extern "C" std::uint32_t rf_build_bvh_device_last_phases(float* outPhaseMs)
{
    if (outPhaseMs) std::memcpy(outPhaseMs, g_lastPhaseMs, sizeof(g_lastPhaseMs));
    return g_lastLevels;
}

#define RF_BUILD_CUDA(expr)                                                                                              \
    do                                                                                                                   \
    {                                                                                                                    \
        const cudaError_t err_ = (expr);                                                                                 \
        if (err_ != cudaSuccess) return setError(RF_ERROR_CUDA, "%s (%s:%d)", cudaGetErrorString(err_), __FILE__, __LINE__); \
    } while (0)

extern "C" rf_status rf_build_bvh_device(
    const rf_positions* triangles,
    const std::uint64_t num_triangles,
    const std::int32_t  device,
    rf_bvh_node*        out_nodes,
    std::uint64_t*      out_num_nodes,
    std::uint64_t*      out_triangle_indices,
    float*              out_device_ms)
{
    if (!triangles || num_triangles == 0 || !out_nodes || !out_num_nodes || !out_triangle_indices)
        return setError(RF_ERROR_INVALID_ARGUMENT, "rf_build_bvh_device: null or empty argument");
    if (num_triangles >= (1ull << 31)) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_build_bvh_device: too many triangles for u32 offsets");
    int deviceCount = 0;
    if (cudaGetDeviceCount(&deviceCount) != cudaSuccess || deviceCount == 0)
    {
        cudaGetLastError();
        return setError(RF_ERROR_CUDA, "No CUDA device available: rf_build_bvh_device has no CPU fallback (rf_build_bvh is the host builder).");
    }
    if (device >= deviceCount) return setError(RF_ERROR_INVALID_ARGUMENT, "CUDA device %d out of range (%d devices).", device, deviceCount);
    if (device >= 0) RF_BUILD_CUDA(cudaSetDevice(device));

    const std::uint32_t n = static_cast<std::uint32_t>(num_triangles);
    const std::uint64_t maxNodes = 2ull * n - 1ull;
    std::lock_guard<std::mutex> lock(g_workspaceMutex);
    int                         currentDevice = 0;
    RF_BUILD_CUDA(cudaGetDevice(&currentDevice));
    BuildWorkspace& ws = g_workspace;
    if (ws.device != currentDevice)
    {
        if (ws.device >= 0 && cudaSetDevice(ws.device) == cudaSuccess) ws.release();
        RF_BUILD_CUDA(cudaSetDevice(currentDevice));
        ws.device = currentDevice;
    }
    auto& dTris = ws.dTris;
    auto& prims = ws.prims;
    auto& order = ws.order;
    auto& owner = ws.owner;
    auto& slotLeft = ws.slotLeft;
    auto& slotRight = ws.slotRight;
    auto& counters = ws.counters;
    auto& nodes = ws.nodes;
    auto& accum = ws.accum;
    auto& buckets = ws.buckets;
    auto& flags = ws.flags;
    auto& scan = ws.scan;
    auto& dIndices = ws.dIndices;
    auto& dOut = ws.dOut;
    auto& scanTemp = ws.scanTemp;
    RF_BUILD_CUDA(dTris.allocate(n));
    RF_BUILD_CUDA(prims.allocate(n));
    RF_BUILD_CUDA(order[0].allocate(n));
    RF_BUILD_CUDA(order[1].allocate(n));
    RF_BUILD_CUDA(owner.allocate(n));
    RF_BUILD_CUDA(slotLeft.allocate(n));
    RF_BUILD_CUDA(slotRight.allocate(n));
    RF_BUILD_CUDA(counters.allocate(2));
    RF_BUILD_CUDA(nodes.allocate(maxNodes));
    RF_BUILD_CUDA(accum.allocate(maxNodes));
    RF_BUILD_CUDA(buckets.allocate(n / 3 + 1)); // SAH nodes of one level hold >= 3 primitives each
    RF_BUILD_CUDA(flags.allocate(n + 1ull));
    RF_BUILD_CUDA(scan.allocate(n + 1ull));
    RF_BUILD_CUDA(dIndices.allocate(n));
    RF_BUILD_CUDA(dOut.allocate(maxNodes));
    RF_BUILD_CUDA(cudaMemcpy(dTris.ptr, triangles, n * sizeof(rf_positions), cudaMemcpyHostToDevice));

    cudaEvent_t evBegin = nullptr, evEnd = nullptr;
    RF_BUILD_CUDA(cudaEventCreate(&evBegin));
    RF_BUILD_CUDA(cudaEventCreate(&evEnd));

    if (!g_levelKernels)
    {
        // one persistent launch: as many blocks as are resident together (the grid barrier needs all of them running)
        int numSms = 0, blocksPerSm = 0;
        RF_BUILD_CUDA(cudaDeviceGetAttribute(&numSms, cudaDevAttrMultiProcessorCount, currentDevice));
        RF_BUILD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, k_bvh_build_fused, BUILD_THREADS, 0));
        if (blocksPerSm < 1) return setError(RF_ERROR_CUDA, "rf_build_bvh_device: the build kernel does not fit an SM");
        const int                grid = numSms * std::min(blocksPerSm, 2);
        if (grid > static_cast<int>(FUSED_MAX_GRID)) return setError(RF_ERROR_CUDA, "rf_build_bvh_device: more resident blocks than the build kernel is laid out for");
        auto& blockTotals = ws.blockTotals;
        auto& levelStartDev = ws.levelStartDev;
        auto& leafStart = ws.leafStart;
        auto& deferList = ws.deferList;
        auto& control = ws.control;
        RF_BUILD_CUDA(blockTotals.allocate(static_cast<std::size_t>(grid)));
        RF_BUILD_CUDA(leafStart.allocate(n + 1ull));
        RF_BUILD_CUDA(deferList.allocate(n));
        RF_BUILD_CUDA(levelStartDev.allocate(FUSED_MAX_LEVELS));
        RF_BUILD_CUDA(control.allocate(1));
        RF_BUILD_CUDA(cudaMemset(control.ptr, 0, sizeof(FusedControl)));
        cudaEvent_t evLocal = nullptr, evNumber = nullptr;
        RF_BUILD_CUDA(cudaEventCreate(&evLocal));
        RF_BUILD_CUDA(cudaEventCreate(&evNumber));
        RF_BUILD_CUDA(cudaEventRecord(evBegin));
        // the grid-wide levels | the deferred subtrees, one block each (compiled with L1 caching) | leaf scan | node records
        k_bvh_build_fused<<<grid, BUILD_THREADS>>>(dTris.ptr, n, prims.ptr, order[0].ptr, order[1].ptr, owner.ptr, slotLeft.ptr, slotRight.ptr, counters.ptr,
                                                   nodes.ptr, accum.ptr, buckets.ptr, flags.ptr, scan.ptr, blockTotals.ptr, levelStartDev.ptr, control.ptr,
                                                   leafStart.ptr, deferList.ptr);
        RF_BUILD_CUDA(cudaEventRecord(evLocal));
        launchBvhBuildLocal(numSms * LOCAL_BLOCKS_PER_SM, nullptr, n, prims.ptr, order[0].ptr, order[1].ptr, owner.ptr, slotLeft.ptr, slotRight.ptr, counters.ptr, nodes.ptr, accum.ptr,
                            buckets.ptr, flags.ptr, scan.ptr, control.ptr, leafStart.ptr, deferList.ptr);
        RF_BUILD_CUDA(cudaEventRecord(evNumber));
        k_bvh_leaf_scan<<<grid, BUILD_THREADS>>>(n, leafStart.ptr, scan.ptr, blockTotals.ptr);
        k_bvh_emit_closed<<<grid, BUILD_THREADS>>>(n, counters.ptr, nodes.ptr, scan.ptr, blockTotals.ptr, order[0].ptr, order[1].ptr, control.ptr, dOut.ptr, dIndices.ptr);
        RF_BUILD_CUDA(cudaEventRecord(evEnd));
        RF_BUILD_CUDA(cudaEventSynchronize(evEnd));
        RF_BUILD_CUDA(cudaGetLastError());
        FusedControl result{};
        RF_BUILD_CUDA(cudaMemcpy(&result, control.ptr, sizeof(result), cudaMemcpyDeviceToHost));
        if (result.error != 0u || result.numNodes > maxNodes) return setError(RF_ERROR_CUDA, "rf_build_bvh_device: the tree has too many levels (internal limit)");
        float fusedMs = 0.f;
        cudaEventElapsedTime(&fusedMs, evBegin, evEnd);
        if (out_device_ms) *out_device_ms = fusedMs;
        for (int k = 0; k < 12; ++k) g_lastPhaseMs[k] = static_cast<float>(result.phaseNs[k]) * 1e-6f;
        cudaEventElapsedTime(&g_lastPhaseMs[11], evLocal, evNumber); // block-local subtrees
        cudaEventElapsedTime(&g_lastPhaseMs[9], evNumber, evEnd);    // leaf scan + node records
        cudaEventDestroy(evBegin), cudaEventDestroy(evEnd), cudaEventDestroy(evLocal), cudaEventDestroy(evNumber);
        g_lastLevels = result.numLevels;
        RF_BUILD_CUDA(cudaMemcpy(out_nodes, dOut.ptr, result.numNodes * sizeof(rf_bvh_node), cudaMemcpyDeviceToHost));
        RF_BUILD_CUDA(cudaMemcpy(out_triangle_indices, dIndices.ptr, n * sizeof(std::uint64_t), cudaMemcpyDeviceToHost));
        *out_num_nodes = result.numNodes;
        return RF_OK;
    }

    // (the level-by-level path scans with CUB; the single-launch build has its own in-place scan)
    std::size_t scanTempBytes = 0;
    RF_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scanTempBytes, flags.ptr, scan.ptr, static_cast<int>(n + 1u)));
    RF_BUILD_CUDA(scanTemp.allocate(scanTempBytes));
    RF_BUILD_CUDA(cudaEventRecord(evBegin));

    k_bvh_prims<<<gridOf(n), BUILD_THREADS>>>(dTris.ptr, n, prims.ptr, order[0].ptr, owner.ptr);
    k_bvh_root<<<1, 1>>>(nodes.ptr, accum.ptr, n);
    const std::uint32_t one[2] = {1u, 0u};
    RF_BUILD_CUDA(cudaMemcpy(counters.ptr, one, sizeof(one), cudaMemcpyHostToDevice));

    std::vector<std::uint32_t> levelStart{0u}; // node slots of level l = [levelStart[l], levelStart[l + 1])
    std::uint32_t              levelBegin = 0, levelEnd = 1;
    int                        cur = 0;
    while (levelBegin != levelEnd)
    {
        const unsigned levelGrid = gridOf(levelEnd - levelBegin);
        RF_BUILD_CUDA(cudaMemsetAsync(counters.ptr + 1, 0, sizeof(std::uint32_t)));
        k_bvh_boxes<<<gridOf(n), BUILD_THREADS>>>(n, prims.ptr, order[cur].ptr, owner.ptr, accum.ptr);
        k_bvh_decide<<<levelGrid, BUILD_THREADS>>>(levelBegin, levelEnd, nodes.ptr, accum.ptr, buckets.ptr, prims.ptr, order[cur].ptr, owner.ptr, counters.ptr);
        k_bvh_buckets<<<gridOf(n), BUILD_THREADS>>>(n, prims.ptr, order[cur].ptr, owner.ptr, nodes.ptr, buckets.ptr);
        k_bvh_sweep<<<levelGrid, BUILD_THREADS>>>(levelBegin, levelEnd, nodes.ptr, accum.ptr, buckets.ptr, owner.ptr, counters.ptr);
        k_bvh_flags<<<gridOf(n + 1ull), BUILD_THREADS>>>(n, prims.ptr, order[cur].ptr, owner.ptr, nodes.ptr, flags.ptr);
        RF_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(scanTemp.ptr, scanTempBytes, flags.ptr, scan.ptr, static_cast<int>(n + 1u)));
        k_bvh_pair<<<gridOf(n), BUILD_THREADS>>>(n, owner.ptr, nodes.ptr, flags.ptr, scan.ptr, slotLeft.ptr, slotRight.ptr);
        k_bvh_permute<<<gridOf(n), BUILD_THREADS>>>(n, owner.ptr, nodes.ptr, flags.ptr, scan.ptr, slotLeft.ptr, slotRight.ptr, order[cur].ptr, order[cur ^ 1].ptr);
        cur ^= 1;
        std::uint32_t created = 0;
        RF_BUILD_CUDA(cudaMemcpy(&created, counters.ptr, sizeof(created), cudaMemcpyDeviceToHost)); // also the level's sync point
        levelStart.push_back(levelEnd);
        levelBegin = levelEnd;
        levelEnd = created;
        if (levelEnd > maxNodes) return setError(RF_ERROR_CUDA, "rf_build_bvh_device: node count overflow (internal error)");
    }
    const std::uint32_t numNodes = levelEnd;
    const int           numLevels = static_cast<int>(levelStart.size()) - 1;
    for (int l = numLevels - 1; l >= 0; --l)
        k_bvh_sizes<<<gridOf(levelStart[l + 1] - levelStart[l]), BUILD_THREADS>>>(levelStart[l], levelStart[l + 1], nodes.ptr);
    for (int l = 0; l < numLevels; ++l)
        k_bvh_preorder<<<gridOf(levelStart[l + 1] - levelStart[l]), BUILD_THREADS>>>(levelStart[l], levelStart[l + 1], nodes.ptr);
    k_bvh_emit<<<gridOf(numNodes), BUILD_THREADS>>>(numNodes, nodes.ptr, dOut.ptr);
    k_bvh_indices<<<gridOf(n), BUILD_THREADS>>>(n, order[cur].ptr, dIndices.ptr);
    RF_BUILD_CUDA(cudaEventRecord(evEnd));
    RF_BUILD_CUDA(cudaEventSynchronize(evEnd));
    RF_BUILD_CUDA(cudaGetLastError());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, evBegin, evEnd);
    cudaEventDestroy(evBegin), cudaEventDestroy(evEnd);
    if (out_device_ms) *out_device_ms = ms;

    RF_BUILD_CUDA(cudaMemcpy(out_nodes, dOut.ptr, numNodes * sizeof(rf_bvh_node), cudaMemcpyDeviceToHost));
    RF_BUILD_CUDA(cudaMemcpy(out_triangle_indices, dIndices.ptr, n * sizeof(std::uint64_t), cudaMemcpyDeviceToHost));
    *out_num_nodes = numNodes;
    return RF_OK;
}

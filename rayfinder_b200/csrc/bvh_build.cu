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
#include "bvh_common.h"
#include "rf_internal.h"

#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

namespace rfb200
{
namespace
{
constexpr std::uint32_t NONE = 0xFFFFFFFFu;
constexpr int           BUILD_THREADS = 256;

enum NodeKind : std::uint32_t
{
    KIND_OPEN = 0,    // created, not decided yet
    KIND_LEAF = 1,
    KIND_SPLIT2 = 2,  // two primitives: median split, done in the decide step
    KIND_SAH = 3,     // waiting for its bucket sweep
    KIND_SPLIT = 4    // SAH split chosen: partition pending / done
};

struct BuildNode
{
    std::uint32_t begin, end;     // primitive positions [begin, end)
    std::uint32_t kind;
    std::uint32_t axis;           // split axis (interior)
    std::uint32_t child0, child1; // node slots
    std::uint32_t bucketSlot;     // SAH: index into the bucket accumulators of the level
    std::uint32_t splitBucket;    // SAH: primitives of buckets <= splitBucket go left
    std::uint32_t mid;            // first position of the second child
    std::uint32_t size;           // nodes in the subtree
    std::uint32_t preorder;       // final node index
    float         cLo, cHi;       // centroid bounds on `axis`
    Box           box;
};

// Reduction state of one node.  Box keys: see loKey() / hiKey().
struct NodeAccum
{
    unsigned long long boxLo[3], boxHi[3];
    std::uint32_t      centLo[3], centHi[3];
};

struct BucketAccum
{
    std::uint32_t count[BVH_NUM_BUCKETS];
    std::uint32_t lo[BVH_NUM_BUCKETS][3], hi[BVH_NUM_BUCKETS][3]; // ordered-uint floats
};

struct Prim
{
    float4 lo; // box.lo.xyz, centroid.x
    float4 hi; // box.hi.xyz, centroid.y
    float  cz; // centroid.z
};

// float <-> uint32 that orders like the float (negative values reversed below the positive ones).
__device__ __forceinline__ std::uint32_t orderedBits(const float f)
{
    const std::uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fromOrderedBits(const std::uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
// Keys of the node-box reduction.  The host folds `lo = (p < lo) ? p : lo` and `hi = (hi < p) ? p : hi` over the
// primitives in sequence order: of numerically equal candidates the earliest wins.  Only -0.0f and +0.0f are equal
// with different bits, so the value part of the key maps both to +0.0f and the position breaks the tie:
//   min key = (ordered(value) << 32) | position             -> atomicMin
//   max key = (ordered(value) << 32) | (0xFFFFFFFF - position) -> atomicMax
__device__ __forceinline__ unsigned long long loKey(const float v, const std::uint32_t pos)
{
    return (static_cast<unsigned long long>(orderedBits(v == 0.0f ? 0.0f : v)) << 32) | pos;
}
__device__ __forceinline__ unsigned long long hiKey(const float v, const std::uint32_t pos)
{
    return (static_cast<unsigned long long>(orderedBits(v == 0.0f ? 0.0f : v)) << 32) | (0xFFFFFFFFu - pos);
}

__device__ __forceinline__ void resetAccum(NodeAccum& a)
{
    for (int k = 0; k < 3; ++k)
    {
        a.boxLo[k] = ~0ull, a.boxHi[k] = 0ull;
        a.centLo[k] = 0xFFFFFFFFu, a.centHi[k] = 0u;
    }
}

__device__ __forceinline__ float comp(const float4 v, const int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ float centroidOf(const Prim& p, const std::uint32_t axis) { return axis == 0 ? p.lo.w : (axis == 1 ? p.hi.w : p.cz); }

// ---- once: primitive boxes (aabb(Positions), aabb.hpp:66-71) and centroids (0.5f * (min + max), aabb.hpp:29) ------
__global__ void k_bvh_prims(const rf_positions* __restrict__ tris, const std::uint32_t n, Prim* prims, std::uint32_t* order, std::uint32_t* owner)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rf_positions t = tris[i];
    const V3           p0 = v3(t.v0), p1 = v3(t.v1), p2 = v3(t.v2);
    const Box          box = makeBox(vmin(vmin(p0, p1), p2), vmax(vmax(p0, p1), p2));
    const V3           c = 0.5f * (box.lo + box.hi);
    prims[i] = Prim{make_float4(box.lo.x, box.lo.y, box.lo.z, c.x), make_float4(box.hi.x, box.hi.y, box.hi.z, c.y), c.z};
    order[i] = i;
    owner[i] = 0u;
}

__global__ void k_bvh_root(BuildNode* nodes, NodeAccum* accum, const std::uint32_t n)
{
    BuildNode root{};
    root.begin = 0u, root.end = n, root.kind = KIND_OPEN, root.child0 = NONE, root.child1 = NONE;
    nodes[0] = root;
    resetAccum(accum[0]);
}

// ---- boxes: fold every primitive of an open node into the node's accumulators ------------------------------------
__global__ void k_bvh_boxes(const std::uint32_t n, const Prim* __restrict__ prims, const std::uint32_t* __restrict__ order,
                            const std::uint32_t* __restrict__ owner, NodeAccum* accum)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const std::uint32_t node = i < n ? owner[i] : NONE;
    const bool          active = node != NONE;
    const unsigned      activeMask = __ballot_sync(0xFFFFFFFFu, active);
    if (!active) return;
    const Prim          p = prims[order[i]];
    const std::uint32_t leader = static_cast<std::uint32_t>(__ffs(static_cast<int>(activeMask)) - 1);
    // (only full warps take the reduced path: the butterfly below needs all 32 lanes)
    const bool          uniform = activeMask == 0xFFFFFFFFu && __all_sync(activeMask, node == __shfl_sync(activeMask, node, leader));
    NodeAccum&          a = accum[node];
    for (int k = 0; k < 3; ++k)
    {
        unsigned long long lo = loKey(comp(p.lo, k), i), hi = hiKey(comp(p.hi, k), i);
        std::uint32_t      cl = orderedBits(centroidOf(p, k)), ch = cl;
        if (uniform)
        {
            // the whole warp folds into one node: reduce first, one atomic per component
            for (int d = 16; d > 0; d >>= 1)
            {
                const unsigned long long lo2 = __shfl_xor_sync(0xFFFFFFFFu, lo, d), hi2 = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
                const std::uint32_t      cl2 = __shfl_xor_sync(0xFFFFFFFFu, cl, d), ch2 = __shfl_xor_sync(0xFFFFFFFFu, ch, d);
                lo = min(lo, lo2), hi = max(hi, hi2), cl = min(cl, cl2), ch = max(ch, ch2);
            }
            if ((threadIdx.x & 31u) != leader) continue;
        }
        atomicMin(&a.boxLo[k], lo);
        atomicMax(&a.boxHi[k], hi);
        atomicMin(&a.centLo[k], cl);
        atomicMax(&a.centHi[k], ch);
    }
}

// ---- decide: leaf / two-primitive median split / SAH (bvh.cpp:96-140) --------------------------------------------
__global__ void k_bvh_decide(const std::uint32_t levelBegin, const std::uint32_t levelEnd, BuildNode* nodes, NodeAccum* accum,
                             BucketAccum* buckets, const Prim* __restrict__ prims, std::uint32_t* order, std::uint32_t* owner,
                             std::uint32_t* counters /* [0] node slots, [1] bucket slots of this level */)
{
    const std::uint32_t s = levelBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelEnd) return;
    BuildNode        nd = nodes[s];
    const NodeAccum& a = accum[s];
    // the winners of the key reductions are positions; the box takes their actual bits (signed zeros included)
    Box nodeBox, centroidBox;
    {
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k)
        {
            const std::uint32_t posLo = static_cast<std::uint32_t>(a.boxLo[k] & 0xFFFFFFFFull);
            const std::uint32_t posHi = 0xFFFFFFFFu - static_cast<std::uint32_t>(a.boxHi[k] & 0xFFFFFFFFull);
            lo[k] = comp(prims[order[posLo]].lo, k);
            hi[k] = comp(prims[order[posHi]].hi, k);
            // Aabb(p1, p2) re-applies min/max at every merge (aabb.hpp:20-26): max = (lo < hi) ? hi : lo, so a box
            // that is flat at zero carries the bits of its lower bound in both
            if (!(lo[k] < hi[k])) hi[k] = lo[k];
        }
        nodeBox.lo = v3(lo[0], lo[1], lo[2]), nodeBox.hi = v3(hi[0], hi[1], hi[2]);
        centroidBox.lo = v3(fromOrderedBits(a.centLo[0]), fromOrderedBits(a.centLo[1]), fromOrderedBits(a.centLo[2]));
        centroidBox.hi = v3(fromOrderedBits(a.centHi[0]), fromOrderedBits(a.centHi[1]), fromOrderedBits(a.centHi[2]));
    }
    nd.box = nodeBox;
    const int           axis = widestAxis(centroidBox);
    const float         cLo = axisOf(centroidBox.lo, axis), cHi = axisOf(centroidBox.hi, axis);
    const std::uint32_t count = nd.end - nd.begin;
    nd.axis = static_cast<std::uint32_t>(axis), nd.cLo = cLo, nd.cHi = cHi;

    if (area(nodeBox) == 0.0f || cLo == cHi || count == 1u)
    {
        nd.kind = KIND_LEAF;
        for (std::uint32_t i = nd.begin; i < nd.end; ++i) owner[i] = NONE; // (leaves with many primitives are rare)
    }
    else if (count < 3u)
    {
        // std::nth_element over two elements: insertion sort, i.e. swap when the second compares less (bvh.cpp:124-137)
        const std::uint32_t p0 = order[nd.begin], p1 = order[nd.begin + 1u];
        if (centroidOf(prims[p1], nd.axis) < centroidOf(prims[p0], nd.axis)) order[nd.begin] = p1, order[nd.begin + 1u] = p0;
        const std::uint32_t c = atomicAdd(&counters[0], 2u);
        nd.kind = KIND_SPLIT2, nd.mid = nd.begin + 1u, nd.child0 = c, nd.child1 = c + 1u;
        BuildNode child{};
        child.kind = KIND_OPEN, child.child0 = NONE, child.child1 = NONE;
        child.begin = nd.begin, child.end = nd.mid;
        nodes[c] = child;
        child.begin = nd.mid, child.end = nd.end;
        nodes[c + 1u] = child;
        resetAccum(accum[c]), resetAccum(accum[c + 1u]);
        owner[nd.begin] = c, owner[nd.begin + 1u] = c + 1u;
    }
    else
    {
        nd.kind = KIND_SAH;
        nd.bucketSlot = atomicAdd(&counters[1], 1u);
        BucketAccum& b = buckets[nd.bucketSlot];
        for (std::size_t k = 0; k < BVH_NUM_BUCKETS; ++k)
        {
            b.count[k] = 0u;
            for (int c = 0; c < 3; ++c) b.lo[k][c] = 0xFFFFFFFFu, b.hi[k][c] = 0u;
        }
    }
    nodes[s] = nd;
}

// ---- buckets: count and bound the primitives of every SAH node per bucket (bvh.cpp:146-156) ----------------------
__global__ void k_bvh_buckets(const std::uint32_t n, const Prim* __restrict__ prims, const std::uint32_t* __restrict__ order,
                              const std::uint32_t* __restrict__ owner, const BuildNode* __restrict__ nodes, BucketAccum* buckets)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const std::uint32_t s = owner[i];
    if (s == NONE) return;
    const BuildNode& nd = nodes[s];
    if (nd.kind != KIND_SAH) return;
    const Prim        p = prims[order[i]];
    const std::size_t b = bvhBucketOf(centroidOf(p, nd.axis), nd.cLo, nd.cHi);
    BucketAccum&      acc = buckets[nd.bucketSlot];
    atomicAdd(&acc.count[b], 1u);
    for (int k = 0; k < 3; ++k)
    {
        atomicMin(&acc.lo[b][k], orderedBits(comp(p.lo, k)));
        atomicMax(&acc.hi[b][k], orderedBits(comp(p.hi, k)));
    }
}

// ---- sweep: the SAH decision of every SAH node (bvh.cpp:157-214), children for the ones that split ---------------
__global__ void k_bvh_sweep(const std::uint32_t levelBegin, const std::uint32_t levelEnd, BuildNode* nodes, NodeAccum* accum,
                            const BucketAccum* __restrict__ buckets, std::uint32_t* owner, std::uint32_t* counters)
{
    const std::uint32_t s = levelBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelEnd) return;
    BuildNode nd = nodes[s];
    if (nd.kind != KIND_SAH) return;
    const BucketAccum& acc = buckets[nd.bucketSlot];
    std::size_t        bucketCount[BVH_NUM_BUCKETS];
    Box                bucketBox[BVH_NUM_BUCKETS];
    for (std::size_t k = 0; k < BVH_NUM_BUCKETS; ++k)
    {
        bucketCount[k] = acc.count[k];
        if (acc.count[k] != 0u)
        {
            // what the host's sequence of grow() leaves in a non-empty bucket: component-wise min / max
            // (signed zeros do not matter here: the boxes only feed area() and comparisons)
            bucketBox[k].lo = v3(fromOrderedBits(acc.lo[k][0]), fromOrderedBits(acc.lo[k][1]), fromOrderedBits(acc.lo[k][2]));
            bucketBox[k].hi = v3(fromOrderedBits(acc.hi[k][0]), fromOrderedBits(acc.hi[k][1]), fromOrderedBits(acc.hi[k][2]));
        }
    }
    const std::uint32_t count = nd.end - nd.begin;
    const int           chosen = bvhChooseSplit(bucketCount, bucketBox, nd.box, count);
    if (chosen < 0)
    {
        nd.kind = KIND_LEAF;
        for (std::uint32_t i = nd.begin; i < nd.end; ++i) owner[i] = NONE; // <= 255 primitives
    }
    else
    {
        std::uint32_t left = 0;
        for (int k = 0; k <= chosen; ++k) left += acc.count[k];
        const std::uint32_t c = atomicAdd(&counters[0], 2u);
        nd.kind = KIND_SPLIT, nd.splitBucket = static_cast<std::uint32_t>(chosen), nd.mid = nd.begin + left, nd.child0 = c, nd.child1 = c + 1u;
        BuildNode child{};
        child.kind = KIND_OPEN, child.child0 = NONE, child.child1 = NONE;
        child.begin = nd.begin, child.end = nd.mid;
        nodes[c] = child;
        child.begin = nd.mid, child.end = nd.end;
        nodes[c + 1u] = child;
        resetAccum(accum[c]), resetAccum(accum[c + 1u]);
    }
    nodes[s] = nd;
}

// ---- partition, step 1: (fails, satisfies) flags of the predicate `bucket <= splitBucket` (bvh.cpp:216-221) -------
__device__ __forceinline__ bool goesLeft(const BuildNode& nd, const Prim& p)
{
    return bvhBucketOf(centroidOf(p, nd.axis), nd.cLo, nd.cHi) <= nd.splitBucket;
}

__global__ void k_bvh_flags(const std::uint32_t n, const Prim* __restrict__ prims, const std::uint32_t* __restrict__ order,
                            const std::uint32_t* __restrict__ owner, const BuildNode* __restrict__ nodes, unsigned long long* flags)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    unsigned long long f = 0ull; // element n: the scan's total
    if (i < n)
    {
        const std::uint32_t s = owner[i];
        if (s != NONE && nodes[s].kind == KIND_SPLIT) f = goesLeft(nodes[s], prims[order[i]]) ? 1ull : (1ull << 32);
    }
    flags[i] = f;
}

// ---- partition, steps 2 and 3.  scan[i] = (fails before i) << 32 | (satisfies before i).  Left zone [begin, mid),
// right zone [mid, end): the k-th failing element of the left zone (from the left) and the k-th satisfying element
// of the right zone (from the right) trade places — libstdc++'s std::__partition for bidirectional iterators. -------
__global__ void k_bvh_pair(const std::uint32_t n, const std::uint32_t* __restrict__ owner, const BuildNode* __restrict__ nodes,
                           const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ scan,
                           std::uint32_t* slotLeft, std::uint32_t* slotRight)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const std::uint32_t s = owner[i];
    if (s == NONE) return;
    const BuildNode& nd = nodes[s];
    if (nd.kind != KIND_SPLIT) return;
    const bool left = flags[i] == 1ull;
    if (i < nd.mid && !left)
    {
        const std::uint32_t k = static_cast<std::uint32_t>((scan[i] >> 32) - (scan[nd.begin] >> 32));
        slotLeft[nd.begin + k] = i;
    }
    else if (i >= nd.mid && left)
    {
        const std::uint32_t k = static_cast<std::uint32_t>((scan[nd.end] & 0xFFFFFFFFull) - (scan[i + 1u] & 0xFFFFFFFFull));
        slotRight[nd.begin + k] = i;
    }
}

__global__ void k_bvh_permute(const std::uint32_t n, std::uint32_t* owner, const BuildNode* __restrict__ nodes,
                              const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ scan,
                              const std::uint32_t* __restrict__ slotLeft, const std::uint32_t* __restrict__ slotRight,
                              const std::uint32_t* __restrict__ orderIn, std::uint32_t* orderOut)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const std::uint32_t s = owner[i];
    std::uint32_t       src = i;
    if (s != NONE && nodes[s].kind == KIND_SPLIT)
    {
        const BuildNode& nd = nodes[s];
        const bool       left = flags[i] == 1ull;
        if (i < nd.mid && !left)
            src = slotRight[nd.begin + static_cast<std::uint32_t>((scan[i] >> 32) - (scan[nd.begin] >> 32))];
        else if (i >= nd.mid && left)
            src = slotLeft[nd.begin + static_cast<std::uint32_t>((scan[nd.end] & 0xFFFFFFFFull) - (scan[i + 1u] & 0xFFFFFFFFull))];
        owner[i] = i < nd.mid ? nd.child0 : nd.child1;
    }
    orderOut[i] = orderIn[src];
}

// ---- numbering: subtree sizes (deepest level first), pre-order indices (root first), node records ------------------
__global__ void k_bvh_sizes(const std::uint32_t levelBegin, const std::uint32_t levelEnd, BuildNode* nodes)
{
    const std::uint32_t s = levelBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelEnd) return;
    BuildNode& nd = nodes[s];
    nd.size = nd.kind == KIND_LEAF ? 1u : 1u + nodes[nd.child0].size + nodes[nd.child1].size;
}

__global__ void k_bvh_preorder(const std::uint32_t levelBegin, const std::uint32_t levelEnd, BuildNode* nodes)
{
    const std::uint32_t s = levelBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= levelEnd) return;
    const BuildNode& nd = nodes[s];
    if (s == 0u) nodes[0].preorder = 0u;
    if (nd.kind == KIND_LEAF) return;
    const std::uint32_t me = s == 0u ? 0u : nd.preorder;
    nodes[nd.child0].preorder = me + 1u;                          // bvh.cpp:93-94: the first child follows its parent
    nodes[nd.child1].preorder = me + 1u + nodes[nd.child0].size;  // the second one follows the first subtree
}

__global__ void k_bvh_emit(const std::uint32_t numNodes, const BuildNode* __restrict__ nodes, rf_bvh_node* out)
{
    const std::uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= numNodes) return;
    const BuildNode& nd = nodes[s];
    rf_bvh_node      o{}; // padding words are zero, as in the reference's aggregate initialisation
    o.aabb_min[0] = nd.box.lo.x, o.aabb_min[1] = nd.box.lo.y, o.aabb_min[2] = nd.box.lo.z;
    o.aabb_max[0] = nd.box.hi.x, o.aabb_max[1] = nd.box.hi.y, o.aabb_max[2] = nd.box.hi.z;
    if (nd.kind == KIND_LEAF)
    {
        o.triangles_offset = nd.begin, o.second_child_offset = 0u, o.triangle_count = nd.end - nd.begin, o.split_axis = 0xFFFFFFFFu; // bvh.cpp:31-42
    }
    else
    {
        o.triangles_offset = 0u, o.second_child_offset = nodes[nd.child1].preorder, o.triangle_count = 0u, o.split_axis = nd.axis; // bvh.cpp:44-55
    }
    out[nd.preorder] = o;
}

__global__ void k_bvh_indices(const std::uint32_t n, const std::uint32_t* __restrict__ order, unsigned long long* triangleIndices)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) triangleIndices[order[i]] = i; // bvh.cpp:64-69: old index -> position in leaf order
}

template<typename T>
struct Buf
{
    T* ptr = nullptr;
    ~Buf() { cudaFree(ptr); }
    cudaError_t allocate(std::size_t count) { return cudaMalloc(&ptr, std::max<std::size_t>(count, 1) * sizeof(T)); }
};
inline unsigned gridOf(std::uint64_t items) { return static_cast<unsigned>((items + BUILD_THREADS - 1) / BUILD_THREADS); }
} // namespace
} // namespace rfb200

using namespace rfb200;

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
    Buf<rf_positions>       dTris;
    Buf<Prim>               prims;
    Buf<std::uint32_t>      order[2], owner, slotLeft, slotRight, counters;
    Buf<BuildNode>          nodes;
    Buf<NodeAccum>          accum;
    Buf<BucketAccum>        buckets;
    Buf<unsigned long long> flags, scan, dIndices;
    Buf<rf_bvh_node>        dOut;
    Buf<unsigned char>      scanTemp;
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
    std::size_t scanTempBytes = 0;
    RF_BUILD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scanTempBytes, flags.ptr, scan.ptr, static_cast<int>(n + 1u)));
    RF_BUILD_CUDA(scanTemp.allocate(scanTempBytes));
    RF_BUILD_CUDA(cudaMemcpy(dTris.ptr, triangles, n * sizeof(rf_positions), cudaMemcpyHostToDevice));

    cudaEvent_t evBegin = nullptr, evEnd = nullptr;
    RF_BUILD_CUDA(cudaEventCreate(&evBegin));
    RF_BUILD_CUDA(cudaEventCreate(&evEnd));
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

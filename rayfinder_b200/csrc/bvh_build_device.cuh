// Device side of the GPU BVH builder: shared by bvh_build.cu (the grid-wide flow, compiled with -Xptxas -dlcm=cg) and
// bvh_build_local.cu (the block-local subtrees, compiled with the default load caching).  See bvh_build.cu for the algorithm.
#pragma once

#include "bvh_common.h"
#include "rf_internal.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
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
    KIND_SPLIT = 4,   // SAH split chosen: partition pending / done
    KIND_DEFERRED = 5 // single-launch build: a small subtree taken out of the level-by-level flow; one block builds it on its own
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
    std::uint32_t size;           // nodes in the subtree            } level-by-level path only
    std::uint32_t preorder;       // final node index                }
    std::uint32_t depth;          // ancestors
    std::uint32_t rightTurns;     // ancestors (incl. the parent) in whose SECOND child the node lies
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

// Where the decide / sweep steps put what they create beyond the node records (all optional).
struct BuildSinks
{
    std::uint32_t* leafStart = nullptr;     // [n + 1]: 1 at the first position of every leaf (closed-form numbering, see preorderOf)
    std::uint32_t* created = nullptr;       // list of the node slots created (a block that builds a subtree on its own collects its next level here)
    std::uint32_t* createdCount = nullptr;
    std::uint32_t* deferList = nullptr;     // decide: nodes with at most deferMaxPrims primitives are not decided but appended here
    std::uint32_t* deferCount = nullptr;
    std::uint32_t  deferMaxPrims = 0;
};

// The exclusive scan of the partition flags as the pair / permute steps read it: plain (level-by-level path), with the offset of
// the slice a position lies in added on the fly (single-launch build: every block scans one slice), or with the value at the end
// of a subtree's range held aside (a block that builds a subtree on its own must not write the position after its range).
struct ScanView
{
    const unsigned long long* scan = nullptr;
    const unsigned long long* slicePrefix = nullptr;
    std::uint32_t             slice = 1;
    std::uint32_t             endPos = 0xFFFFFFFFu;
    unsigned long long        endValue = 0;
    __device__ __forceinline__ unsigned long long at(const std::uint32_t i) const
    {
        if (i == endPos) return endValue;
        return scan[i] + (slicePrefix ? slicePrefix[i / slice] : 0ull);
    }
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
// (called by whole warps: lane L handles position i = warpBase + L)
__device__ __forceinline__ void boxesAt(const std::uint32_t i, const std::uint32_t n, const Prim* __restrict__ prims, const std::uint32_t* __restrict__ order,
                                        const std::uint32_t* __restrict__ owner, NodeAccum* __restrict__ accum)
{
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
__global__ void k_bvh_boxes(const std::uint32_t n, const Prim* __restrict__ prims, const std::uint32_t* __restrict__ order,
                            const std::uint32_t* __restrict__ owner, NodeAccum* accum)
{
    boxesAt(blockIdx.x * blockDim.x + threadIdx.x, n, prims, order, owner, accum);
}

// ---- decide: leaf / two-primitive median split / SAH (bvh.cpp:96-140) --------------------------------------------
__device__ __forceinline__ void decideAt(const std::uint32_t s, BuildNode* nodes, NodeAccum* accum, BucketAccum* buckets, const Prim* __restrict__ prims,
                                         std::uint32_t* order, std::uint32_t* owner, std::uint32_t* nodeCounter /* node slots */,
                                         std::uint32_t* bucketCounter /* bucket slots of this level */, const BuildSinks sinks = BuildSinks{})
{
    BuildNode        nd = nodes[s];
    const NodeAccum& a = accum[s];
    // the winners of the key reductions are positions; the box takes their actual bits (signed zeros included)
    Box nodeBox, centroidBox;
    {
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k)
        {
            // (the accumulators are updated by atomics, i.e. in L2: read them there — k_bvh_build_local runs with L1 caching)
            const std::uint32_t posLo = static_cast<std::uint32_t>(__ldcg(&a.boxLo[k]) & 0xFFFFFFFFull);
            const std::uint32_t posHi = 0xFFFFFFFFu - static_cast<std::uint32_t>(__ldcg(&a.boxHi[k]) & 0xFFFFFFFFull);
            lo[k] = comp(prims[order[posLo]].lo, k);
            hi[k] = comp(prims[order[posHi]].hi, k);
            // Aabb(p1, p2) re-applies min/max at every merge (aabb.hpp:20-26): max = (lo < hi) ? hi : lo, so a box
            // that is flat at zero carries the bits of its lower bound in both
            if (!(lo[k] < hi[k])) hi[k] = lo[k];
        }
        nodeBox.lo = v3(lo[0], lo[1], lo[2]), nodeBox.hi = v3(hi[0], hi[1], hi[2]);
        centroidBox.lo = v3(fromOrderedBits(__ldcg(&a.centLo[0])), fromOrderedBits(__ldcg(&a.centLo[1])), fromOrderedBits(__ldcg(&a.centLo[2])));
        centroidBox.hi = v3(fromOrderedBits(__ldcg(&a.centHi[0])), fromOrderedBits(__ldcg(&a.centHi[1])), fromOrderedBits(__ldcg(&a.centHi[2])));
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
        if (sinks.leafStart) sinks.leafStart[nd.begin] = 1u;
    }
    else if (sinks.deferList && count <= sinks.deferMaxPrims)
    {
        // a small subtree: its primitives leave the level-by-level flow, one block will build it on its own
        nd.kind = KIND_DEFERRED;
        sinks.deferList[atomicAdd(sinks.deferCount, 1u)] = s;
        for (std::uint32_t i = nd.begin; i < nd.end; ++i) owner[i] = NONE;
    }
    else if (count < 3u)
    {
        // std::nth_element over two elements: insertion sort, i.e. swap when the second compares less (bvh.cpp:124-137)
        const std::uint32_t p0 = order[nd.begin], p1 = order[nd.begin + 1u];
        if (centroidOf(prims[p1], nd.axis) < centroidOf(prims[p0], nd.axis)) order[nd.begin] = p1, order[nd.begin + 1u] = p0;
        const std::uint32_t c = atomicAdd(nodeCounter, 2u);
        nd.kind = KIND_SPLIT2, nd.mid = nd.begin + 1u, nd.child0 = c, nd.child1 = c + 1u;
        BuildNode child{};
        child.kind = KIND_OPEN, child.child0 = NONE, child.child1 = NONE;
        child.depth = nd.depth + 1u;
        child.begin = nd.begin, child.end = nd.mid, child.rightTurns = nd.rightTurns;
        nodes[c] = child;
        child.begin = nd.mid, child.end = nd.end, child.rightTurns = nd.rightTurns + 1u;
        nodes[c + 1u] = child;
        resetAccum(accum[c]), resetAccum(accum[c + 1u]);
        owner[nd.begin] = c, owner[nd.begin + 1u] = c + 1u;
        if (sinks.created)
        {
            const std::uint32_t at = atomicAdd(sinks.createdCount, 2u);
            sinks.created[at] = c, sinks.created[at + 1u] = c + 1u;
        }
    }
    else
    {
        nd.kind = KIND_SAH;
        nd.bucketSlot = atomicAdd(bucketCounter, 1u);
        BucketAccum& b = buckets[nd.bucketSlot];
        for (std::size_t k = 0; k < BVH_NUM_BUCKETS; ++k)
        {
            b.count[k] = 0u;
            for (int c = 0; c < 3; ++c) b.lo[k][c] = 0xFFFFFFFFu, b.hi[k][c] = 0u;
        }
    }
    nodes[s] = nd;
}
__global__ void k_bvh_decide(const std::uint32_t levelBegin, const std::uint32_t levelEnd, BuildNode* nodes, NodeAccum* accum,
                             BucketAccum* buckets, const Prim* __restrict__ prims, std::uint32_t* order, std::uint32_t* owner, std::uint32_t* counters)
{
    const std::uint32_t s = levelBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s < levelEnd) decideAt(s, nodes, accum, buckets, prims, order, owner, &counters[0], &counters[1]);
}

// ---- buckets: count and bound the primitives of every SAH node per bucket (bvh.cpp:146-156) ----------------------
__device__ __forceinline__ void bucketsAt(const std::uint32_t i, const Prim* __restrict__ prims, const std::uint32_t* __restrict__ order,
                                          const std::uint32_t* __restrict__ owner, const BuildNode* __restrict__ nodes, BucketAccum* __restrict__ buckets)
{
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
__global__ void k_bvh_buckets(const std::uint32_t n, const Prim* __restrict__ prims, const std::uint32_t* __restrict__ order,
                              const std::uint32_t* __restrict__ owner, const BuildNode* __restrict__ nodes, BucketAccum* buckets)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bucketsAt(i, prims, order, owner, nodes, buckets);
}

// ---- sweep: the SAH decision of every SAH node (bvh.cpp:157-214), children for the ones that split ---------------
__device__ __forceinline__ void sweepAt(const std::uint32_t s, BuildNode* nodes, NodeAccum* accum, const BucketAccum* buckets, std::uint32_t* owner,
                                        std::uint32_t* nodeCounter, const BuildSinks sinks = BuildSinks{})
{
    BuildNode nd = nodes[s];
    if (nd.kind != KIND_SAH) return;
    const BucketAccum& acc = buckets[nd.bucketSlot];
    std::size_t        bucketCount[BVH_NUM_BUCKETS];
    Box                bucketBox[BVH_NUM_BUCKETS];
    for (std::size_t k = 0; k < BVH_NUM_BUCKETS; ++k)
    {
        bucketCount[k] = __ldcg(&acc.count[k]); // (accumulated by atomics in L2, see decideAt)
        if (bucketCount[k] != 0u)
        {
            // what the host's sequence of grow() leaves in a non-empty bucket: component-wise min / max
            // (signed zeros do not matter here: the boxes only feed area() and comparisons)
            bucketBox[k].lo = v3(fromOrderedBits(__ldcg(&acc.lo[k][0])), fromOrderedBits(__ldcg(&acc.lo[k][1])), fromOrderedBits(__ldcg(&acc.lo[k][2])));
            bucketBox[k].hi = v3(fromOrderedBits(__ldcg(&acc.hi[k][0])), fromOrderedBits(__ldcg(&acc.hi[k][1])), fromOrderedBits(__ldcg(&acc.hi[k][2])));
        }
    }
    const std::uint32_t count = nd.end - nd.begin;
    const int           chosen = bvhChooseSplit(bucketCount, bucketBox, nd.box, count);
    if (chosen < 0)
    {
        nd.kind = KIND_LEAF;
        for (std::uint32_t i = nd.begin; i < nd.end; ++i) owner[i] = NONE; // <= 255 primitives
        if (sinks.leafStart) sinks.leafStart[nd.begin] = 1u;
    }
    else
    {
        std::uint32_t left = 0;
        for (int k = 0; k <= chosen; ++k) left += static_cast<std::uint32_t>(bucketCount[k]);
        const std::uint32_t c = atomicAdd(nodeCounter, 2u);
        nd.kind = KIND_SPLIT, nd.splitBucket = static_cast<std::uint32_t>(chosen), nd.mid = nd.begin + left, nd.child0 = c, nd.child1 = c + 1u;
        BuildNode child{};
        child.kind = KIND_OPEN, child.child0 = NONE, child.child1 = NONE;
        child.depth = nd.depth + 1u;
        child.begin = nd.begin, child.end = nd.mid, child.rightTurns = nd.rightTurns;
        nodes[c] = child;
        child.begin = nd.mid, child.end = nd.end, child.rightTurns = nd.rightTurns + 1u;
        nodes[c + 1u] = child;
        resetAccum(accum[c]), resetAccum(accum[c + 1u]);
        if (sinks.created)
        {
            const std::uint32_t at = atomicAdd(sinks.createdCount, 2u);
            sinks.created[at] = c, sinks.created[at + 1u] = c + 1u;
        }
    }
    nodes[s] = nd;
}
__global__ void k_bvh_sweep(const std::uint32_t levelBegin, const std::uint32_t levelEnd, BuildNode* nodes, NodeAccum* accum,
                            const BucketAccum* __restrict__ buckets, std::uint32_t* owner, std::uint32_t* counters)
{
    const std::uint32_t s = levelBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s < levelEnd) sweepAt(s, nodes, accum, buckets, owner, &counters[0]);
}

// ---- partition, step 1: (fails, satisfies) flags of the predicate `bucket <= splitBucket` (bvh.cpp:216-221) -------
__device__ __forceinline__ bool goesLeft(const BuildNode& nd, const Prim& p)
{
    return bvhBucketOf(centroidOf(p, nd.axis), nd.cLo, nd.cHi) <= nd.splitBucket;
}

__device__ __forceinline__ unsigned long long flagAt(const std::uint32_t i, const std::uint32_t n, const Prim* __restrict__ prims,
                                                     const std::uint32_t* order, const std::uint32_t* owner, const BuildNode* nodes)
{
    unsigned long long f = 0ull; // element n: the scan's total
    if (i < n)
    {
        const std::uint32_t s = owner[i];
        if (s != NONE && nodes[s].kind == KIND_SPLIT) f = goesLeft(nodes[s], prims[order[i]]) ? 1ull : (1ull << 32);
    }
    return f;
}
__global__ void k_bvh_flags(const std::uint32_t n, const Prim* __restrict__ prims, const std::uint32_t* __restrict__ order,
                            const std::uint32_t* __restrict__ owner, const BuildNode* __restrict__ nodes, unsigned long long* flags)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) flags[i] = flagAt(i, n, prims, order, owner, nodes);
}

// ---- partition, steps 2 and 3.  scan[i] = (fails before i) << 32 | (satisfies before i).  Left zone [begin, mid),
// right zone [mid, end): the k-th failing element of the left zone (from the left) and the k-th satisfying element
// of the right zone (from the right) trade places — libstdc++'s std::__partition for bidirectional iterators. -------
__device__ __forceinline__ void pairAt(const std::uint32_t i, const std::uint32_t* owner, const BuildNode* nodes, const unsigned long long* flags,
                                       const ScanView scan, std::uint32_t* slotLeft, std::uint32_t* slotRight)
{
    const std::uint32_t s = owner[i];
    if (s == NONE) return;
    const BuildNode& nd = nodes[s];
    if (nd.kind != KIND_SPLIT) return;
    const bool left = flags[i] == 1ull;
    if (i < nd.mid && !left)
    {
        const std::uint32_t k = static_cast<std::uint32_t>((scan.at(i) >> 32) - (scan.at(nd.begin) >> 32));
        slotLeft[nd.begin + k] = i;
    }
    else if (i >= nd.mid && left)
    {
        const std::uint32_t k = static_cast<std::uint32_t>((scan.at(nd.end) & 0xFFFFFFFFull) - (scan.at(i + 1u) & 0xFFFFFFFFull));
        slotRight[nd.begin + k] = i;
    }
}
__global__ void k_bvh_pair(const std::uint32_t n, const std::uint32_t* __restrict__ owner, const BuildNode* __restrict__ nodes,
                           const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ scan,
                           std::uint32_t* slotLeft, std::uint32_t* slotRight)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pairAt(i, owner, nodes, flags, ScanView{scan}, slotLeft, slotRight);
}

__device__ __forceinline__ void permuteAt(const std::uint32_t i, std::uint32_t* owner, const BuildNode* nodes, const unsigned long long* flags,
                                          const ScanView scan, const std::uint32_t* slotLeft, const std::uint32_t* slotRight,
                                          const std::uint32_t* orderIn, std::uint32_t* orderOut)
{
    const std::uint32_t s = owner[i];
    std::uint32_t       src = i;
    if (s != NONE && nodes[s].kind == KIND_SPLIT)
    {
        const BuildNode& nd = nodes[s];
        const bool       left = flags[i] == 1ull;
        if (i < nd.mid && !left)
            src = slotRight[nd.begin + static_cast<std::uint32_t>((scan.at(i) >> 32) - (scan.at(nd.begin) >> 32))];
        else if (i >= nd.mid && left)
            src = slotLeft[nd.begin + static_cast<std::uint32_t>((scan.at(nd.end) & 0xFFFFFFFFull) - (scan.at(i + 1u) & 0xFFFFFFFFull))];
        owner[i] = i < nd.mid ? nd.child0 : nd.child1;
    }
    orderOut[i] = orderIn[src];
}
__global__ void k_bvh_permute(const std::uint32_t n, std::uint32_t* owner, const BuildNode* __restrict__ nodes,
                              const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ scan,
                              const std::uint32_t* __restrict__ slotLeft, const std::uint32_t* __restrict__ slotRight,
                              const std::uint32_t* __restrict__ orderIn, std::uint32_t* orderOut)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) permuteAt(i, owner, nodes, flags, ScanView{scan}, slotLeft, slotRight, orderIn, orderOut);
}

// ---- numbering: subtree sizes (deepest level first), pre-order indices (root first), node records ------------------
__device__ __forceinline__ void sizeAt(const std::uint32_t s, BuildNode* nodes)
{
    BuildNode& nd = nodes[s];
    nd.size = nd.kind == KIND_LEAF ? 1u : 1u + nodes[nd.child0].size + nodes[nd.child1].size;
}
__global__ void k_bvh_sizes(const std::uint32_t levelBegin, const std::uint32_t levelEnd, BuildNode* nodes)
{
    const std::uint32_t s = levelBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s < levelEnd) sizeAt(s, nodes);
}

__device__ __forceinline__ void preorderAt(const std::uint32_t s, BuildNode* nodes)
{
    const BuildNode& nd = nodes[s];
    if (s == 0u) nodes[0].preorder = 0u;
    if (nd.kind == KIND_LEAF) return;
    const std::uint32_t me = s == 0u ? 0u : nd.preorder;
    nodes[nd.child0].preorder = me + 1u;                          // bvh.cpp:93-94: the first child follows its parent
    nodes[nd.child1].preorder = me + 1u + nodes[nd.child0].size;  // the second one follows the first subtree
}
__global__ void k_bvh_preorder(const std::uint32_t levelBegin, const std::uint32_t levelEnd, BuildNode* nodes)
{
    const std::uint32_t s = levelBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s < levelEnd) preorderAt(s, nodes);
}

// Pre-order (depth-first) index of a node in closed form.  The nodes before X in pre-order are its ancestors and the complete
// subtrees hanging to the LEFT of the path root -> X, one per ancestor in whose second child X lies.  Those subtrees hold exactly
// the leaves that start before X.begin — L of them —, and a binary subtree with l leaves has 2 l - 1 nodes, so
//     preorder(X) = depth(X) + 2 L(X.begin) - rightTurns(X):
// no pass over the levels (bottom-up sizes, top-down indices: 2 x 31 grid barriers for Sponza), just one scan of the leaf starts.
__device__ __forceinline__ std::uint32_t preorderOf(const BuildNode& nd, const ScanView leavesBefore)
{
    return nd.depth + 2u * static_cast<std::uint32_t>(leavesBefore.at(nd.begin)) - nd.rightTurns;
}
__device__ __forceinline__ void emitClosedForm(const std::uint32_t s, const BuildNode* nodes, const ScanView leavesBefore, rf_bvh_node* out)
{
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
        o.triangles_offset = 0u, o.second_child_offset = preorderOf(nodes[nd.child1], leavesBefore), o.triangle_count = 0u, o.split_axis = nd.axis; // bvh.cpp:44-55
    }
    out[preorderOf(nd, leavesBefore)] = o;
}

__device__ __forceinline__ void emitAt(const std::uint32_t s, const BuildNode* nodes, rf_bvh_node* out)
{
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
__global__ void k_bvh_emit(const std::uint32_t numNodes, const BuildNode* __restrict__ nodes, rf_bvh_node* out)
{
    const std::uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < numNodes) emitAt(s, nodes, out);
}

__global__ void k_bvh_indices(const std::uint32_t n, const std::uint32_t* __restrict__ order, unsigned long long* triangleIndices)
{
    const std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) triangleIndices[order[i]] = i; // bvh.cpp:64-69: old index -> position in leaf order
}

// warp min / max of 64-bit keys over the lanes of `mask` (all lanes of `mask` call)
__device__ __forceinline__ unsigned long long __reduce_min_sync_u64(const unsigned mask, unsigned long long v)
{
    for (int d = 16; d > 0; d >>= 1)
    {
        const unsigned long long other = __shfl_xor_sync(mask, v, d);
        const bool               valid = (mask >> ((threadIdx.x & 31u) ^ static_cast<unsigned>(d))) & 1u;
        if (valid && other < v) v = other;
    }
    return v;
}
__device__ __forceinline__ unsigned long long __reduce_max_sync_u64(const unsigned mask, unsigned long long v)
{
    for (int d = 16; d > 0; d >>= 1)
    {
        const unsigned long long other = __shfl_xor_sync(mask, v, d);
        const bool               valid = (mask >> ((threadIdx.x & 31u) ^ static_cast<unsigned>(d))) & 1u;
        if (valid && other > v) v = other;
    }
    return v;
}

// ---- the whole build as ONE persistent launch ---------------------------------------------------------------------
// The level-by-level path above costs ~10 launches and one host read-back per level (~50 levels for Sponza: 5.4 ms, nearly
// all of it launch latency and synchronisation).  Here every phase is a grid-stride loop of the same device functions inside
// one kernel whose blocks are all resident, separated by a grid-wide barrier (~2 us instead of a launch boundary), and the
// level bookkeeping stays on the device.  The partition's scan is done in place: every block scans its contiguous slice of
// the flags; the slice totals (one per block) are scanned by each block for itself and added on the fly (ScanView).
//
// Three things keep the number of grid barriers down (Sponza, 31 levels: 3.0 -> see DESIGN.md):
//   * 7 barriers per level (boxes | decide | buckets | sweep | scan | pair | permute): the slice offsets and the level
//     bookkeeping need none of their own;
//   * a node with at most SUBTREE_MAX_PRIMS primitives leaves the level-by-level flow when it is decided (KIND_DEFERRED); after
//     the last level ONE BLOCK builds each of those subtrees completely on its own, with the same device functions on the
//     same global arrays restricted to the subtree's positions, separated by __syncthreads — the bottom half of the levels
//     costs no grid barrier at all.  Measured alternatives: one WARP per subtree of <= 256 primitives, all subtrees at once
//     (2.55 ms for Sponza against 2.26 ms: three more grid-wide levels, and a lane walks 8 positions per step one after the
//     other); 2 048 primitives per block (2.69 ms: a block's time grows with the positions per thread, 0.58 ms for 1 024
//     primitives, 1.58 ms for 2 048).  A level of a small subtree is seven dependent steps of 1.5-6 us each through L2; staging
//     a subtree's positions in shared memory is what is left;
//   * node numbers come from a closed form (preorderOf) instead of a bottom-up and a top-down pass over the levels.
constexpr std::uint32_t SUBTREE_MAX_PRIMS = 1024;
constexpr std::uint32_t FUSED_MAX_GRID = 1024;
#ifndef RF_BVH_LOCAL_BLOCKS
#define RF_BVH_LOCAL_BLOCKS 2
#endif
constexpr int           LOCAL_BLOCKS_PER_SM = RF_BVH_LOCAL_BLOCKS; // blocks per SM of k_bvh_build_local (no grid barrier there: any grid works)

struct FusedControl
{
    unsigned int  barrier;      // grid barrier: arrivals so far (monotonic)
    std::uint32_t numLevels;    // levels built by the whole grid
    std::uint32_t numNodes;
    std::uint32_t error;        // 1: more than MAX_LEVELS levels
    std::uint32_t deferCount;   // subtrees handed to single blocks
    std::uint32_t deferCursor[2];
    std::uint32_t finalOrderIsSecond; // which of the two order buffers holds the positions after the last grid-wide level
    unsigned long long phaseNs[12]; // time block 0 spent in each phase incl. its barrier (diagnostics): boxes, decide, buckets, sweep, scan, -, pair,
                                    // permute, -, leaf scan, emit, block-local subtrees
};
constexpr std::uint32_t FUSED_MAX_LEVELS = 4096;

__device__ __forceinline__ void gridBarrier(FusedControl* ctl, unsigned int& generation)
{
    __syncthreads();
    ++generation;
    if (threadIdx.x == 0)
    {
        __threadfence(); // this block's writes before its arrival
        atomicAdd(&ctl->barrier, 1u);
        const unsigned int target = generation * gridDim.x;
        while (*reinterpret_cast<volatile unsigned int*>(&ctl->barrier) < target) {}
        __threadfence(); // (also invalidates this SM's L1: the other blocks' writes are read from L2)
    }
    __syncthreads();
}

// Exclusive scan over the positions [begin, end) by one block: value(i) -> scanOut[i] (relative to `begin`), and flagsOut[i] =
// value(i) if flagsOut is given.  Returns the total (to every thread).  Both 32-bit halves of a value stay below 2^32.
template<class Value>
__device__ __forceinline__ unsigned long long blockScanRange(const std::uint32_t begin, const std::uint32_t end, const Value value, unsigned long long* flagsOut,
                                                             unsigned long long* scanOut, unsigned long long* warpSums, unsigned long long* carry)
{
    if (threadIdx.x == 0) *carry = 0ull;
    __syncthreads();
    for (std::uint32_t base = begin; base < end; base += BUILD_THREADS)
    {
        const std::uint32_t      i = base + threadIdx.x;
        const unsigned long long f = i < end ? value(i) : 0ull;
        unsigned long long       incl = f;
        for (int d = 1; d < 32; d <<= 1)
        {
            const unsigned long long up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((threadIdx.x & 31u) >= static_cast<unsigned>(d)) incl += up;
        }
        if ((threadIdx.x & 31u) == 31u) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long before = *carry;
        for (std::uint32_t wIdx = 0; wIdx < (threadIdx.x >> 5); ++wIdx) before += warpSums[wIdx];
        if (i < end)
        {
            if (flagsOut) flagsOut[i] = f;
            scanOut[i] = before + incl - f;
        }
        __syncthreads();
        if (threadIdx.x == BUILD_THREADS - 1) *carry = before + incl;
        __syncthreads();
    }
    return *carry;
}

__global__ void __launch_bounds__(BUILD_THREADS, 2) k_bvh_build_fused(
    const rf_positions* __restrict__ tris, const std::uint32_t n, Prim* prims, std::uint32_t* order0, std::uint32_t* order1, std::uint32_t* owner,
    std::uint32_t* slotLeft, std::uint32_t* slotRight, std::uint32_t* counters, BuildNode* nodes, NodeAccum* accum, BucketAccum* buckets,
    unsigned long long* flags, unsigned long long* scan, unsigned long long* blockTotals, std::uint32_t* levelStart, FusedControl* ctl,
    std::uint32_t* leafStart, std::uint32_t* deferList)
{
    __shared__ unsigned long long warpSums[BUILD_THREADS / 32];
    __shared__ unsigned long long sliceCarry;
    __shared__ unsigned long long slicePrefix[FUSED_MAX_GRID]; // exclusive scan of the blocks' slice totals
    __shared__ BucketAccum        blockBuckets;
    __shared__ NodeAccum          blockAccum;
    unsigned int        generation = 0;
    const std::uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    unsigned long long  phaseStart = 0;
    const auto          now = []() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    };
    const auto endPhase = [&](const int phase) { // barrier + bookkeeping of the phase's duration as block 0 sees it
        gridBarrier(ctl, generation);
        if (tid == 0)
        {
            const unsigned long long t = now();
            ctl->phaseNs[phase] += t - phaseStart;
            phaseStart = t;
        }
    };
    // the blocks' slice totals (blockTotals, complete after a grid barrier) -> their exclusive scan in shared memory
    const auto scanSliceTotals = [&]() {
        blockScanRange(0u, gridDim.x, [&](const std::uint32_t b) { return blockTotals[b]; }, nullptr, slicePrefix, warpSums, &sliceCarry);
    };
    const std::uint32_t nPadded = (n + 31u) & ~31u;

    // primitives and the root (k_bvh_prims, k_bvh_root)
    for (std::uint32_t i = tid; i < n; i += stride)
    {
        const rf_positions t = tris[i];
        const V3           p0 = v3(t.v0), p1 = v3(t.v1), p2 = v3(t.v2);
        const Box          box = makeBox(vmin(vmin(p0, p1), p2), vmax(vmax(p0, p1), p2));
        const V3           c = 0.5f * (box.lo + box.hi);
        prims[i] = Prim{make_float4(box.lo.x, box.lo.y, box.lo.z, c.x), make_float4(box.hi.x, box.hi.y, box.hi.z, c.y), c.z};
        order0[i] = i;
        owner[i] = 0u;
    }
    for (std::uint32_t i = tid; i <= n; i += stride) leafStart[i] = 0u;
    if (tid == 0)
    {
        BuildNode root{};
        root.begin = 0u, root.end = n, root.kind = KIND_OPEN, root.child0 = NONE, root.child1 = NONE;
        nodes[0] = root;
        resetAccum(accum[0]);
        counters[0] = 1u, counters[1] = 0u;
        levelStart[0] = 0u;
    }
    gridBarrier(ctl, generation);
    if (tid == 0) phaseStart = now();

    std::uint32_t  levelBegin = 0, levelEnd = 1, level = 0;
    std::uint32_t* order = order0;
    std::uint32_t* orderNext = order1;
    // contiguous slice of the positions [0, n] (n + 1 flags: the last one is the scan's total) every block scans
    const std::uint32_t slice = (n + 1u + gridDim.x - 1u) / gridDim.x;
    const std::uint32_t sliceBegin = min(blockIdx.x * slice, n + 1u), sliceEnd = min(sliceBegin + slice, n + 1u);
    BuildSinks          levelSinks;
    levelSinks.leafStart = leafStart, levelSinks.deferList = deferList, levelSinks.deferCount = &ctl->deferCount, levelSinks.deferMaxPrims = SUBTREE_MAX_PRIMS;
    while (levelBegin != levelEnd)
    {
        // boxes (same remark as for the buckets below: a round that lies in one node is folded in shared memory first)
        for (std::uint32_t base = blockIdx.x * BUILD_THREADS; base < nPadded; base += stride)
        {
            const std::uint32_t i = base + threadIdx.x;
            const std::uint32_t node = i < n ? owner[i] : NONE;
            const std::uint32_t first = owner[min(base, n - 1u)];
            const int           uniform = __syncthreads_and((i >= n || node == first) ? 1 : 0) && first != NONE;
            if (!uniform)
            {
                boxesAt(i, n, prims, order, owner, accum);
                continue;
            }
            if (threadIdx.x == 0) resetAccum(blockAccum);
            __syncthreads();
            if (i < n)
            {
                const Prim p = prims[order[i]];
                for (int k = 0; k < 3; ++k)
                {
                    unsigned long long lo = loKey(comp(p.lo, k), i), hi = hiKey(comp(p.hi, k), i);
                    std::uint32_t      cl = orderedBits(centroidOf(p, k)), ch = cl;
                    const unsigned     mask = __activemask();
                    lo = __reduce_min_sync_u64(mask, lo), hi = __reduce_max_sync_u64(mask, hi);
                    cl = __reduce_min_sync(mask, cl), ch = __reduce_max_sync(mask, ch);
                    if ((threadIdx.x & 31u) == static_cast<unsigned>(__ffs(static_cast<int>(mask)) - 1))
                    {
                        atomicMin(&blockAccum.boxLo[k], lo);
                        atomicMax(&blockAccum.boxHi[k], hi);
                        atomicMin(&blockAccum.centLo[k], cl);
                        atomicMax(&blockAccum.centHi[k], ch);
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x < 3)
            {
                NodeAccum& a = accum[first];
                atomicMin(&a.boxLo[threadIdx.x], blockAccum.boxLo[threadIdx.x]);
                atomicMax(&a.boxHi[threadIdx.x], blockAccum.boxHi[threadIdx.x]);
                atomicMin(&a.centLo[threadIdx.x], blockAccum.centLo[threadIdx.x]);
                atomicMax(&a.centHi[threadIdx.x], blockAccum.centHi[threadIdx.x]);
            }
            __syncthreads();
        }
        endPhase(0);
        for (std::uint32_t s = levelBegin + tid; s < levelEnd; s += stride) decideAt(s, nodes, accum, buckets, prims, order, owner, &counters[0], &counters[1], levelSinks);
        endPhase(1);
        // buckets.  The 256 consecutive positions a block handles per round mostly lie in ONE node while nodes are large, and
        // then all 256 threads would hammer the same 84 accumulator words in L2 (measured: 2.4 of 5.3 ms for Sponza): such a
        // round is folded in shared memory first and leaves with one atomic per word.
        for (std::uint32_t base = blockIdx.x * BUILD_THREADS; base < n; base += stride)
        {
            const std::uint32_t i = base + threadIdx.x;
            const std::uint32_t node = i < n ? owner[i] : NONE;
            const std::uint32_t first = owner[base];
            const bool          sahHere = node != NONE && nodes[node].kind == KIND_SAH;
            const int           uniform = __syncthreads_and((i >= n || node == first) ? 1 : 0) && first != NONE;
            if (!uniform)
            {
                if (i < n) bucketsAt(i, prims, order, owner, nodes, buckets);
                continue;
            }
            if (nodes[first].kind != KIND_SAH) continue; // (uniform for the block: `first` is)
            for (std::uint32_t k = threadIdx.x; k < BVH_NUM_BUCKETS; k += BUILD_THREADS)
            {
                blockBuckets.count[k] = 0u;
                for (int c = 0; c < 3; ++c) blockBuckets.lo[k][c] = 0xFFFFFFFFu, blockBuckets.hi[k][c] = 0u;
            }
            __syncthreads();
            if (sahHere)
            {
                const BuildNode&  nd = nodes[node];
                const Prim        p = prims[order[i]];
                const std::size_t b = bvhBucketOf(centroidOf(p, nd.axis), nd.cLo, nd.cHi);
                atomicAdd(&blockBuckets.count[b], 1u);
                for (int k = 0; k < 3; ++k)
                {
                    atomicMin(&blockBuckets.lo[b][k], orderedBits(comp(p.lo, k)));
                    atomicMax(&blockBuckets.hi[b][k], orderedBits(comp(p.hi, k)));
                }
            }
            __syncthreads();
            BucketAccum& acc = buckets[nodes[first].bucketSlot];
            for (std::uint32_t k = threadIdx.x; k < BVH_NUM_BUCKETS; k += BUILD_THREADS)
            {
                if (blockBuckets.count[k] == 0u) continue;
                atomicAdd(&acc.count[k], blockBuckets.count[k]);
                for (int c = 0; c < 3; ++c)
                {
                    atomicMin(&acc.lo[k][c], blockBuckets.lo[k][c]);
                    atomicMax(&acc.hi[k][c], blockBuckets.hi[k][c]);
                }
            }
            __syncthreads();
        }
        endPhase(2);
        for (std::uint32_t s = levelBegin + tid; s < levelEnd; s += stride) sweepAt(s, nodes, accum, buckets, owner, &counters[0], levelSinks);
        endPhase(3);
        // flags + exclusive scan of this block's slice (relative to the slice), slice total
        {
            const unsigned long long total = blockScanRange(
                sliceBegin, sliceEnd, [&](const std::uint32_t i) { return flagAt(i, n, prims, order, owner, nodes); }, flags, scan, warpSums, &sliceCarry);
            if (threadIdx.x == 0) blockTotals[blockIdx.x] = total;
        }
        endPhase(4);
        scanSliceTotals(); // (offset of a slice = totals of the slices before it, added on the fly by ScanView::at)
        ScanView view;
        view.scan = scan, view.slicePrefix = slicePrefix, view.slice = slice;
        for (std::uint32_t i = tid; i < n; i += stride) pairAt(i, owner, nodes, flags, view, slotLeft, slotRight);
        endPhase(6);
        for (std::uint32_t i = tid; i < n; i += stride) permuteAt(i, owner, nodes, flags, view, slotLeft, slotRight, order, orderNext);
        {
            std::uint32_t* t = order;
            order = orderNext, orderNext = t;
        }
        endPhase(7);
        // next level: the node slots created by this one.  (No barrier of its own: the counter changes again in the next level's
        // decide step, i.e. after the next barrier, which every block reaches only after it has read the counter here.)
        const std::uint32_t created = *reinterpret_cast<volatile std::uint32_t*>(&counters[0]);
        ++level;
        if (tid == 0)
        {
            if (level < FUSED_MAX_LEVELS) levelStart[level] = levelEnd;
            counters[1] = 0u; // bucket slots of the next level
        }
        levelBegin = levelEnd, levelEnd = created;
        if (level >= FUSED_MAX_LEVELS - 1u)
        {
            if (tid == 0) ctl->error = 1u;
            break;
        }
    }
    if (tid == 0) ctl->numLevels = level, ctl->finalOrderIsSecond = order == order1 ? 1u : 0u;
}

// ---- after the grid-wide levels: the deferred subtrees, one block each.  A kernel of its own (bvh_build_local.cu) because it
// is compiled with the default load caching: everything a block reads here except the atomically updated accumulators (read
// with ld.global.cg in decideAt / sweepAt) was written by the block itself, i.e. through the same SM's L1, so the dependent
// loads of a step hit L1 instead of making two or three L2 round trips each (the grid-wide flow above passes data between SMs
// inside one launch and has to bypass L1: -Xptxas -dlcm=cg for bvh_build.cu).
__global__ void __launch_bounds__(BUILD_THREADS, LOCAL_BLOCKS_PER_SM) k_bvh_build_local(
    const std::uint32_t n, const Prim* __restrict__ prims, std::uint32_t* order0, std::uint32_t* order1, std::uint32_t* owner, std::uint32_t* slotLeft,
    std::uint32_t* slotRight, std::uint32_t* counters, BuildNode* nodes, NodeAccum* accum, BucketAccum* buckets, unsigned long long* flags,
    unsigned long long* scan, FusedControl* ctl, std::uint32_t* leafStart, const std::uint32_t* __restrict__ deferList)
{
    __shared__ unsigned long long warpSums[BUILD_THREADS / 32];
    __shared__ unsigned long long sliceCarry;
    // node slots of the current / next level of the subtree, their counts, the level's bucket slots
    __shared__ std::uint32_t localLevel[2][SUBTREE_MAX_PRIMS];
    __shared__ std::uint32_t localCount[2], localBucketCount, localSubtree;
    std::uint32_t* const order = ctl->finalOrderIsSecond ? order1 : order0;
    std::uint32_t* const orderNext = ctl->finalOrderIsSecond ? order0 : order1;
    (void)n;
    // ---- the deferred subtrees, one block each --------------------------------------------------------------------------
    {
        // (two passes over the list, the larger subtrees first: the blocks then finish closer together)
        const std::uint32_t deferred = *reinterpret_cast<volatile std::uint32_t*>(&ctl->deferCount);
        std::uint32_t       pass = 0;
        while (true)
        {
            if (threadIdx.x == 0) localSubtree = atomicAdd(&ctl->deferCursor[pass], 1u);
            __syncthreads();
            const std::uint32_t k = localSubtree;
            __syncthreads();
            if (k >= deferred)
            {
                if (++pass == 2u) break;
                continue;
            }
            const std::uint32_t root = deferList[k];
            const std::uint32_t b = nodes[root].begin, e = nodes[root].end;
            if (((e - b) > SUBTREE_MAX_PRIMS / 2u) != (pass == 0u)) continue; // not this pass's size class
            // SAH nodes of one level of this subtree hold >= 3 primitives each: slots [b / 3, e / 3) of the level-by-level flow's
            // bucket accumulators (idle now) are this subtree's own
            BucketAccum* const myBuckets = buckets + b / 3u;
            __syncthreads();
            for (std::uint32_t i = b + threadIdx.x; i < e; i += BUILD_THREADS) owner[i] = root;
            if (threadIdx.x == 0)
            {
                nodes[root].kind = KIND_OPEN;
                resetAccum(accum[root]);
                localLevel[0][0] = root;
                localCount[0] = 1u, localCount[1] = 0u, localBucketCount = 0u;
            }
            __syncthreads();
            std::uint32_t* oc = order;
            std::uint32_t* oo = orderNext;
            int            cur = 0;
            while (localCount[cur] != 0u)
            {
                const std::uint32_t levelNodes = localCount[cur];
                BuildSinks          sinks;
                sinks.leafStart = leafStart, sinks.created = localLevel[cur ^ 1], sinks.createdCount = &localCount[cur ^ 1];
                for (std::uint32_t base = b; base < e; base += BUILD_THREADS) boxesAt(base + threadIdx.x, e, prims, oc, owner, accum);
                __syncthreads();
                for (std::uint32_t idx = threadIdx.x; idx < levelNodes; idx += BUILD_THREADS)
                    decideAt(localLevel[cur][idx], nodes, accum, myBuckets, prims, oc, owner, &counters[0], &localBucketCount, sinks);
                __syncthreads();
                for (std::uint32_t i = b + threadIdx.x; i < e; i += BUILD_THREADS) bucketsAt(i, prims, oc, owner, nodes, myBuckets);
                __syncthreads();
                for (std::uint32_t idx = threadIdx.x; idx < levelNodes; idx += BUILD_THREADS)
                    sweepAt(localLevel[cur][idx], nodes, accum, myBuckets, owner, &counters[0], sinks);
                __syncthreads();
                ScanView view;
                view.scan = scan, view.endPos = e;
                view.endValue = blockScanRange(b, e, [&](const std::uint32_t i) { return flagAt(i, e, prims, oc, owner, nodes); }, flags, scan, warpSums, &sliceCarry);
                __syncthreads();
                for (std::uint32_t i = b + threadIdx.x; i < e; i += BUILD_THREADS) pairAt(i, owner, nodes, flags, view, slotLeft, slotRight);
                __syncthreads();
                for (std::uint32_t i = b + threadIdx.x; i < e; i += BUILD_THREADS) permuteAt(i, owner, nodes, flags, view, slotLeft, slotRight, oc, oo);
                {
                    std::uint32_t* t = oc;
                    oc = oo, oo = t;
                }
                __syncthreads();
                if (threadIdx.x == 0) localCount[cur] = 0u, localBucketCount = 0u;
                cur ^= 1;
                __syncthreads();
            }
            if (oc != order)
                for (std::uint32_t i = b + threadIdx.x; i < e; i += BUILD_THREADS) order[i] = oc[i];
            __syncthreads();
        }
    }
}

// ---- numbering: one scan of the leaf starts (preorderOf), then the records ---------------------------------------------
__global__ void __launch_bounds__(BUILD_THREADS) k_bvh_leaf_scan(const std::uint32_t n, const std::uint32_t* __restrict__ leafStart, unsigned long long* scan,
                                                                unsigned long long* blockTotals)
{
    __shared__ unsigned long long warpSums[BUILD_THREADS / 32];
    __shared__ unsigned long long sliceCarry;
    const std::uint32_t slice = (n + 1u + gridDim.x - 1u) / gridDim.x;
    const std::uint32_t sliceBegin = min(blockIdx.x * slice, n + 1u), sliceEnd = min(sliceBegin + slice, n + 1u);
    const unsigned long long total =
        blockScanRange(sliceBegin, sliceEnd, [&](const std::uint32_t i) { return static_cast<unsigned long long>(leafStart[i]); }, nullptr, scan, warpSums, &sliceCarry);
    if (threadIdx.x == 0) blockTotals[blockIdx.x] = total;
}
// (launched with the grid of k_bvh_leaf_scan)
__global__ void __launch_bounds__(BUILD_THREADS) k_bvh_emit_closed(
    const std::uint32_t n, const std::uint32_t* __restrict__ counters, const BuildNode* __restrict__ nodes, const unsigned long long* __restrict__ scan,
    const unsigned long long* __restrict__ blockTotals, const std::uint32_t* order0, const std::uint32_t* order1, FusedControl* ctl, rf_bvh_node* out,
    unsigned long long* triangleIndices)
{
    __shared__ unsigned long long warpSums[BUILD_THREADS / 32];
    __shared__ unsigned long long sliceCarry;
    __shared__ unsigned long long slicePrefix[FUSED_MAX_GRID];
    blockScanRange(0u, gridDim.x, [&](const std::uint32_t b) { return blockTotals[b]; }, nullptr, slicePrefix, warpSums, &sliceCarry);
    const std::uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const std::uint32_t numNodes = counters[0];
    if (tid == 0) ctl->numNodes = numNodes;
    ScanView leavesBefore;
    leavesBefore.scan = scan, leavesBefore.slicePrefix = slicePrefix, leavesBefore.slice = (n + 1u + gridDim.x - 1u) / gridDim.x;
    const std::uint32_t* order = ctl->finalOrderIsSecond ? order1 : order0;
    for (std::uint32_t s = tid; s < numNodes; s += stride) emitClosedForm(s, nodes, leavesBefore, out);
    for (std::uint32_t i = tid; i < n; i += stride) triangleIndices[order[i]] = i; // bvh.cpp:64-69
}


} // namespace

// bvh_build_local.cu: launches k_bvh_build_local as that translation unit compiled it (the pointers are the builder's arrays:
// prims, order0, order1, owner, slotLeft, slotRight, counters, nodes, accum, buckets, flags, scan, control, leafStart, deferList).
void launchBvhBuildLocal(int grid, cudaStream_t stream, std::uint32_t n, const void* prims, void* order0, void* order1, void* owner, void* slotLeft, void* slotRight,
                         void* counters, void* nodes, void* accum, void* buckets, void* flags, void* scan, void* control, void* leafStart, const void* deferList);
} // namespace rfb200

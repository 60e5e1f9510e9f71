// Pieces of buildBvh (common/bvh.cpp:81-291, common/aabb.hpp) shared by the host builder (host_core.cpp) and the
// device builder (bvh_build.cu): the same fp32 expressions on both sides, so the two produce the same bytes.
#pragma once

#include "rf_vec.h"

#include <cfloat>
#include <cstddef>
#include <cstdint>

namespace rfb200
{
struct Box
{
    V3 lo{FLT_MAX, FLT_MAX, FLT_MAX};
    V3 hi{-FLT_MAX, -FLT_MAX, -FLT_MAX};
};
// Aabb(p1, p2) re-applies min/max (aabb.hpp:20-26); merge() goes through it (aabb.hpp:50-58).
RF_HD Box makeBox(V3 a, V3 b) { return Box{vmin(a, b), vmax(a, b)}; }
RF_HD Box grow(const Box& b, V3 p) { return makeBox(vmin(b.lo, p), vmax(b.hi, p)); }
RF_HD Box grow(const Box& a, const Box& b) { return makeBox(vmin(a.lo, b.lo), vmax(a.hi, b.hi)); }
RF_HD float area(const Box& b)
{
    const V3 d = b.hi - b.lo;
    return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z); // aabb.hpp:60-64
}
RF_HD int widestAxis(const Box& b)
{
    const V3 d = b.hi - b.lo; // aabb.hpp:33-48: ties fall through to z
    if (d.x > d.y && d.x > d.z) return 0;
    if (d.y > d.z) return 1;
    return 2;
}
RF_HD float axisOf(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

constexpr std::size_t BVH_NUM_BUCKETS = 12; // bvh.cpp:142-145
constexpr std::size_t BVH_MAX_LEAF = 255;
constexpr float       BVH_TRAVERSAL_COST = 0.5f;
constexpr float       BVH_INTERSECTION_COST = 1.0f;

// size_t(numBuckets * (c - lo) / (hi - lo)), clamped (bvh.cpp:152-155).
RF_HD std::size_t bvhBucketOf(float centroid, float lo, float hi)
{
#if defined(__CUDA_ARCH__)
    const float q = __fdiv_rn(static_cast<float>(BVH_NUM_BUCKETS) * (centroid - lo), hi - lo);
#else
    const float q = static_cast<float>(BVH_NUM_BUCKETS) * (centroid - lo) / (hi - lo);
#endif
    const std::size_t b = static_cast<std::size_t>(q);
    return b < BVH_NUM_BUCKETS - 1 ? b : BVH_NUM_BUCKETS - 1;
}

// The SAH sweep over the 12 buckets (bvh.cpp:157-214): returns the bucket after which to split, or -1 for a
// leaf.  `count` = primitives of the node.
RF_HD int bvhChooseSplit(const std::size_t* bucketCount, const Box* bucketBox, const Box& nodeBox, std::size_t count)
{
    constexpr std::size_t NUM_SPLITS = BVH_NUM_BUCKETS - 1;
    float                 cost[NUM_SPLITS] = {};
    {
        std::size_t below = 0;
        Box         boxBelow;
        for (std::size_t i = 0; i < NUM_SPLITS; ++i)
        {
            below += bucketCount[i];
            boxBelow = grow(boxBelow, bucketBox[i]);
            cost[i] += BVH_INTERSECTION_COST * static_cast<float>(below) * area(boxBelow);
        }
        std::size_t above = 0;
        Box         boxAbove;
        for (std::size_t i = NUM_SPLITS; i > 0; --i)
        {
            above += bucketCount[i];
            boxAbove = grow(boxAbove, bucketBox[i]);
            cost[i - 1] += BVH_INTERSECTION_COST * static_cast<float>(above) * area(boxAbove);
        }
    }
    float minCost = FLT_MAX;
    int   splitBucket = -1;
    for (std::size_t i = 0; i < NUM_SPLITS; ++i)
    {
        if (cost[i] < minCost)
        {
            minCost = cost[i];
            splitBucket = static_cast<int>(i);
        }
    }
    const float leafCost = BVH_INTERSECTION_COST * static_cast<float>(count);
#if defined(__CUDA_ARCH__)
    const float totalCost = BVH_TRAVERSAL_COST + __fdiv_rn(minCost, area(nodeBox));
#else
    const float totalCost = BVH_TRAVERSAL_COST + minCost / area(nodeBox);
#endif
    if (!(count > BVH_MAX_LEAF || totalCost < leafCost)) return -1;
    return splitBucket;
}
} // namespace rfb200

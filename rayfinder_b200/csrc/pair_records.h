// Child-pair records: the traversal layout of the BVH (traversal_pairs.cuh).
//
// The reference visits a node, tests ITS box, and only then learns where to go (common/ray_intersection.cpp:156-204,
// reference_path_tracer.wgsl:376-428): one dependent memory round trip per visit.  Here every interior node has one
// 64-byte record that holds the boxes of BOTH its children, so one load decides two visits:
//
//   * the near child's visit happens right away (nothing can change tmax in between, so the slab test is the one the
//     reference would do);
//   * the far child's slab test is done too, but only its tmax-INDEPENDENT part counts now: the entry pushed for it carries
//     `t` = its slab entry distance tmin (or +inf when the box is missed whatever tmax is), and the reference's result
//     `hit = ... && tmin < rayTMax` (ray_intersection.cpp:135) is completed when the entry is popped, with the tmax of
//     that moment.  A popped entry that misses costs no memory access at all.
//
// Visits, their order, every fp32 operation of the slab test and therefore nodesVisited are unchanged; what changes is that
// a ray needs one record per interior node it ENTERS (about half its visits) instead of one node per visit.
//
// Record r describes interior node n_r (records are numbered in node order, leaves have none):
//   box0 = box of the first child  (node n_r + 1),             link0 = where it leads
//   box1 = box of the second child (node secondChildOffset),   link1 = where it leads
//   meta = split axis of n_r (bits 0-1: which child is near for a ray, bvh.cpp:44-55 / ray_intersection.cpp:184-199)
// A link is either an interior child's record index (bit 31 clear) or a leaf: bit 31 set, bits 30..24 = triangleCount - 1,
// bits 23..0 = trianglesOffset.  Scenes whose leaves do not fit that (>= 2^24 triangles or a leaf with more than 128) are
// traced with the one-node-per-visit kernel (traversal.cuh) instead.
#pragma once

#include "rf_internal.h"

#include <cstdint>
#include <vector>

namespace rfb200
{
struct alignas(64) PairRecord
{
    float         box0[6]; // min.xyz, max.xyz of the first child
    float         box1[6]; // min.xyz, max.xyz of the second child
    std::uint32_t link0, link1;
    std::uint32_t meta;    // split axis of the node itself
    std::uint32_t node;    // the node's index in the reference array (diagnostics only)
};
static_assert(sizeof(PairRecord) == 64, "two 32-byte sectors, one 128-byte line holds two records");

constexpr std::uint32_t PAIR_LINK_LEAF = 0x80000000u;
constexpr std::uint32_t PAIR_LEAF_MAX_TRIANGLES = 128u;
constexpr std::uint32_t PAIR_LEAF_MAX_OFFSET = 1u << 24;

inline bool pairLinkIsLeaf(std::uint32_t link) { return (link & PAIR_LINK_LEAF) != 0u; }

struct PairScene
{
    std::vector<PairRecord> records;  // one per interior node, in node order
    float                   rootBox[6]; // box of node 0 (tested by the ray's first visit)
    std::uint32_t           rootLink;   // record 0, or the leaf link of a single-leaf tree
    bool                    usable = false; // false: some leaf does not fit a link, use the per-node layout
};

// Host-side construction from the reference's node array (already validated: validateBvh in device.cu).
inline PairScene buildPairRecords(const rf_bvh_node* nodes, std::uint64_t numNodes)
{
    PairScene scene;
    for (int a = 0; a < 3; ++a) scene.rootBox[a] = nodes[0].aabb_min[a], scene.rootBox[3 + a] = nodes[0].aabb_max[a];
    // record index of every interior node
    std::vector<std::uint32_t> recordOf(numNodes, 0u);
    std::uint32_t              numRecords = 0;
    for (std::uint64_t i = 0; i < numNodes; ++i)
    {
        const rf_bvh_node& n = nodes[i];
        if (n.triangle_count == 0u)
            recordOf[i] = numRecords++;
        else if (n.triangle_count > PAIR_LEAF_MAX_TRIANGLES || n.triangles_offset >= PAIR_LEAF_MAX_OFFSET)
            return scene; // not usable
    }
    if (numRecords >= (1u << 30)) return scene;
    const auto linkOf = [&](std::uint64_t i) -> std::uint32_t {
        const rf_bvh_node& n = nodes[i];
        if (n.triangle_count == 0u) return recordOf[i];
        return PAIR_LINK_LEAF | ((n.triangle_count - 1u) << 24) | n.triangles_offset;
    };
    scene.records.resize(numRecords);
    for (std::uint64_t i = 0; i < numNodes; ++i)
    {
        const rf_bvh_node& n = nodes[i];
        if (n.triangle_count != 0u) continue;
        PairRecord&        r = scene.records[recordOf[i]];
        const rf_bvh_node& c0 = nodes[i + 1];
        const rf_bvh_node& c1 = nodes[n.second_child_offset];
        for (int a = 0; a < 3; ++a)
        {
            r.box0[a] = c0.aabb_min[a], r.box0[3 + a] = c0.aabb_max[a];
            r.box1[a] = c1.aabb_min[a], r.box1[3 + a] = c1.aabb_max[a];
        }
        r.link0 = linkOf(i + 1);
        r.link1 = linkOf(n.second_child_offset);
        r.meta = n.split_axis;
        r.node = static_cast<std::uint32_t>(i);
    }
    scene.rootLink = linkOf(0);
    scene.usable = true;
    return scene;
}
} // namespace rfb200

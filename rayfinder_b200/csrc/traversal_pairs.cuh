// BVH traversal over child-pair records (pair_records.h) for sm_100a — the device twin of
//   rayIntersectBvh   reference_path_tracer.wgsl:371-429  / common/ray_intersection.cpp:138-213
//   shadowRay         reference_path_tracer.wgsl:323-368
// with the same contract as traversal.cuh (traceRays): per ray, the visit ORDER, every fp32 operation of the slab and
// triangle tests, the hit and nodesVisited are the reference's.  What differs from traceRays is the memory shape of a ray:
//
//   * one 64-byte record per interior node the ray ENTERS (two LDG.256 of one 128-byte line) instead of one 32-byte node per
//     VISIT: the record holds both children's boxes, the near child's visit is decided on the spot and the far child's entry
//     on the stack carries its slab entry distance, so the reference's `tmin < rayTMax` is completed at the pop with the tmax
//     of that moment — a popped entry that misses costs no memory access.  Measured on the rays of the benchmark frame
//     (tools/model/pair_model.cpp): 88.8 visits per ray = 88.8 node loads before, 45.6 record loads now; 18 of the 42 pops per
//     ray end at the compare.
//   * stack entries are (link, t) pairs.  The first PAIR_STACK_SHARED of them live in shared memory (column per thread, bank =
//     lane), deeper ones in a per-thread local array: 99.8 % of the benchmark's rays never go deeper than 16.
//
// The persistent-warp loop around it (per-lane refill, parked triangle rounds, warp votes) is traceRays's.
#pragma once

#include "pair_records.h"
#include "traversal.cuh"

namespace rfb200
{
constexpr int PAIR_STACK_SHARED = 16; // stack entries kept in shared memory (2 words each)

struct PairSceneDevice
{
    const PairRecord* records;
    float             rootBox[6];
    std::uint32_t     rootLink;
};

// Both halves of a record: bytes 0-31 (box0, box1.min.xy) and 32-63 (box1.min.z, box1.max, link0, link1, meta, node).
struct PairHalves
{
    float         a[8];
    float         b[4];
    std::uint32_t link0, link1, meta, node;
};
__device__ __forceinline__ PairHalves loadPair(const PairRecord* p)
{
    PairHalves h;
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(h.a[0]), "=f"(h.a[1]), "=f"(h.a[2]), "=f"(h.a[3]), "=f"(h.a[4]), "=f"(h.a[5]), "=f"(h.a[6]), "=f"(h.a[7])
        : "l"(p));
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8+32];"
        : "=f"(h.b[0]), "=f"(h.b[1]), "=f"(h.b[2]), "=f"(h.b[3]), "=r"(h.link0), "=r"(h.link1), "=r"(h.meta), "=r"(h.node)
        : "l"(p));
    return h;
}

// rayIntersectAabb split at its only tmax-dependent term: returns `t` such that the reference's result for this box is
// `t < rayTMax` — the slab entry distance tmin when the rest of the test passes, +inf when it fails whatever rayTMax is.
// Fast form and NaN guard as in traceRays (traversal.cuh, DESIGN.md "Slab test").
__device__ __noinline__ float slabEntryExact(
    const float minX, const float minY, const float minZ, const float maxX, const float maxY, const float maxZ, const std::uint32_t negMask,
    const V3 o, const float ix, const float iy, const float iz)
{
    // ray_intersection.cpp:101-136, literally; `tmin < rayTMax` left to the caller
    const bool  negX = negMask & 1u, negY = negMask & 2u, negZ = negMask & 4u;
    const float loX = negX ? maxX : minX, hiX = negX ? minX : maxX;
    const float loY = negY ? maxY : minY, hiY = negY ? minY : maxY;
    const float loZ = negZ ? maxZ : minZ, hiZ = negZ ? minZ : maxZ;
    float       tmin = (loX - o.x) * ix;
    float       tmx = (hiX - o.x) * ix;
    const float tymin = (loY - o.y) * iy;
    const float tymax = (hiY - o.y) * iy;
    bool        boxHit = !((tmin > tymax) || (tymin > tmx));
    tmin = (tymin < tmin) ? tmin : tymin; // std::max(tymin, tmin)
    tmx = (tmx < tymax) ? tmx : tymax;    // std::min(tymax, tmax)
    const float tzmin = (loZ - o.z) * iz;
    const float tzmax = (hiZ - o.z) * iz;
    boxHit = boxHit && !((tmin > tzmax) || (tzmin > tmx));
    tmin = (tzmin < tmin) ? tmin : tzmin;
    tmx = (tmx < tzmax) ? tmx : tzmax;
    // a NaN tmin makes `tmin < rayTMax` false, as in the reference; a NaN tmx makes `tmx > 0` false
    return (boxHit && (tmx > 0.0f)) ? tmin : __int_as_float(0x7F800000);
}

__device__ __forceinline__ float slabEntry(
    const float minX, const float minY, const float minZ, const float maxX, const float maxY, const float maxZ, const std::uint32_t negMask,
    const V3 o, const float ix, const float iy, const float iz, const bool exact)
{
    const float x0 = (minX - o.x) * ix, x1 = (maxX - o.x) * ix;
    const float y0 = (minY - o.y) * iy, y1 = (maxY - o.y) * iy;
    const float z0 = (minZ - o.z) * iz, z1 = (maxZ - o.z) * iz;
    const float tmin = max3Nan(minNan(x0, x1), minNan(y0, y1), minNan(z0, z1));
    const float tmx = min3Nan(maxNan(x0, x1), maxNan(y0, y1), maxNan(z0, z1));
    float       t = ((tmin <= tmx) && (tmx > 0.0f)) ? tmin : __int_as_float(0x7F800000);
    if (exact || eitherNan(tmin, tmx)) t = slabEntryExact(minX, minY, minZ, maxX, maxY, maxZ, negMask, o, ix, iy, iz);
    return t;
}

// The persistent traversal loop over pair records.  IO as for traceRays (acquire / fetch / finish); MODE 0 closest hit,
// 1 any-hit, 2 per ray.  VARIANT bits 0-1: rounds per warp vote minus 1; bit 2: closest-hit rays do not push far children
// whose box is missed whatever tmax is (their visit is counted at once: every pushed entry of a closest-hit ray is popped
// eventually, so the total is the reference's; an any-hit ray may stop early and must count each visit when it happens).
//
// SIMT shape of a round (what the lanes of a warp execute together):
//   pops     a short divergent loop: every lane that needs an entry pops until one is entered (two shared-memory reads and a
//            compare per pop — a popped entry that misses costs nothing else);
//   expand   straight-line, predicated: every lane that entered an interior node loads its record, runs both slab tests,
//            pushes the far child and steps to the near one.
// Keeping the expensive part convergent is what the split buys: in a first version with one pop and one expand per lane per
// round the warps ran at 14-15 of 32 lanes per instruction (profiles/r02_pairs_v1_ncu.txt).
constexpr int PAIR_DEFAULT_VARIANT = 5;
template<int MODE, int VARIANT, int BLOCK, class IO>
__device__ __forceinline__ void traceRaysPairs(
    const PairSceneDevice& pairs,
    const float4* __restrict__ tris,
    const bool        sceneOrdered,
    const TraceTuning tuning,
    IO&               io,
    std::uint32_t&    recordsLoaded) // += records this lane loaded (the kernel's memory work, for the roofline)
{
    __shared__ std::uint32_t stackMem[2 * PAIR_STACK_SHARED * BLOCK];
    constexpr std::uint32_t  ENTRY_STRIDE = 2u * BLOCK * 4u; // bytes between consecutive entries of one thread
    constexpr std::uint32_t  WORD_STRIDE = BLOCK * 4u;       // link at +0, t at +WORD_STRIDE
    const std::uint32_t      stackBase = static_cast<std::uint32_t>(__cvta_generic_to_shared(stackMem + threadIdx.x));
    const std::uint32_t      stackLimit = stackBase + PAIR_STACK_SHARED * ENTRY_STRIDE;
    std::uint32_t            stackTop = stackBase; // address the next entry would get if every entry lived in shared memory
    std::uint32_t            deepLink[RF_STACK_SIZE - PAIR_STACK_SHARED]; // entries PAIR_STACK_SHARED.. (rare; local memory)
    float                    deepT[RF_STACK_SIZE - PAIR_STACK_SHARED];

    enum : int
    {
        IDLE = 0,   // no ray
        EXPAND = 1, // next action: load record `cur`, visit its near child, push its far child
        TRI = 2,    // parked at the leaf `cur` (a leaf link): its triangles from number `triDone` on are still to test
        DONE = 3,   // traversal finished, result not yet handed to IO
        POP = 4     // next action: pop an entry and complete its slab test
    };
    constexpr int  ROUNDS_PER_VOTE = (VARIANT & 3) + 1;
    constexpr bool SKIP_DEFINITE_MISSES = (VARIANT & 4) != 0;
    const float    INF = __int_as_float(0x7F800000);

    int           state = IDLE;
    std::uint32_t rayIdx = 0;
    V3            o = v3(0.f, 0.f, 0.f), d = o;
    float         ix = 0.f, iy = 0.f, iz = 0.f, tmax = 0.f;
    std::uint32_t negMask = 0; // bit a = invDir[a] < 0 (dirNeg); bit 3 = literal slab test for every box of this ray
    std::uint32_t cur = 0;     // EXPAND: record index; TRI: the leaf's link
    std::uint32_t triDone = 0, rayNodes = 0, rayTris = 0;
    HitRecord     hit{RF_NO_HIT, 0.f, 0.f, 0.f};
    bool          exhausted = false;
    bool          laneAnyHit = false;
#define RF_ANY_HIT (MODE == 2 ? laneAnyHit : (MODE == 1))

    // Where `link` leads once its box is known to be entered: EXPAND for a record index, TRI for a leaf link (bit 31).
    const auto enter = [&](const std::uint32_t link) {
        cur = link;
        state = EXPAND + static_cast<int>(link >> 31);
    };

    // One pop (ray_intersection.cpp:200-203 followed by the next iteration's :158-160): the entry's visit is counted, its
    // slab test completed against the current tmax.
    const auto popStep = [&]() {
        if (stackTop == stackBase)
        {
            state = DONE;
            return;
        }
        stackTop -= ENTRY_STRIDE;
        std::uint32_t link;
        float         t;
        if (__builtin_expect(stackTop >= stackLimit, 0))
        {
            const std::uint32_t k = (stackTop - stackLimit) / ENTRY_STRIDE;
            link = deepLink[k], t = deepT[k];
        }
        else
        {
            link = stackLoad(stackTop);
            t = __uint_as_float(stackLoad(stackTop + WORD_STRIDE));
        }
        ++rayNodes;
        if (t < tmax) enter(link); // else: missed, stay in POP
    };

    // One interior node the ray has entered: both children's slab tests from one record.
    const auto expandStep = [&]() {
        ++recordsLoaded;
        const PairHalves r = loadPair(pairs.records + cur);
        // fast form of both slab tests (traversal.cuh, DESIGN.md "Slab test"), split at the tmax-dependent term
        const float ax0 = (r.a[0] - o.x) * ix, ax1 = (r.a[3] - o.x) * ix;
        const float ay0 = (r.a[1] - o.y) * iy, ay1 = (r.a[4] - o.y) * iy;
        const float az0 = (r.a[2] - o.z) * iz, az1 = (r.a[5] - o.z) * iz;
        const float bx0 = (r.a[6] - o.x) * ix, bx1 = (r.b[1] - o.x) * ix;
        const float by0 = (r.a[7] - o.y) * iy, by1 = (r.b[2] - o.y) * iy;
        const float bz0 = (r.b[0] - o.z) * iz, bz1 = (r.b[3] - o.z) * iz;
        const float aMin = max3Nan(minNan(ax0, ax1), minNan(ay0, ay1), minNan(az0, az1));
        const float aMax = min3Nan(maxNan(ax0, ax1), maxNan(ay0, ay1), maxNan(az0, az1));
        const float bMin = max3Nan(minNan(bx0, bx1), minNan(by0, by1), minNan(bz0, bz1));
        const float bMax = min3Nan(maxNan(bx0, bx1), maxNan(by0, by1), maxNan(bz0, bz1));
        float       t0 = ((aMin <= aMax) && (aMax > 0.0f)) ? aMin : INF;
        float       t1 = ((bMin <= bMax) && (bMax > 0.0f)) ? bMin : INF;
        if (__builtin_expect((negMask & 8u) != 0u || eitherNan(aMin, aMax) || eitherNan(bMin, bMax), 0))
        {
            // a NaN slab product (0 * inf), or a ray / scene that needs the literal form throughout
            t0 = slabEntryExact(r.a[0], r.a[1], r.a[2], r.a[3], r.a[4], r.a[5], negMask, o, ix, iy, iz);
            t1 = slabEntryExact(r.a[6], r.a[7], r.b[0], r.b[1], r.b[2], r.b[3], negMask, o, ix, iy, iz);
        }
        // near child first by the sign of invDir[splitAxis] (ray_intersection.cpp:184-199); the other one is pushed
        const bool          nearIsSecond = ((negMask >> (r.meta & 3u)) & 1u) != 0u;
        const float         tNear = nearIsSecond ? t1 : t0, tFar = nearIsSecond ? t0 : t1;
        const std::uint32_t linkNear = nearIsSecond ? r.link1 : r.link0, linkFar = nearIsSecond ? r.link0 : r.link1;
        rayNodes += 1; // the near child's visit
        if (SKIP_DEFINITE_MISSES && !RF_ANY_HIT && tFar == INF)
        {
            rayNodes += 1; // the far child's visit will miss whatever happens: count it now, push nothing
        }
        else
        {
            if (__builtin_expect(stackTop >= stackLimit, 0))
            {
                const std::uint32_t k = (stackTop - stackLimit) / ENTRY_STRIDE;
                deepLink[k] = linkFar, deepT[k] = tFar;
            }
            else
            {
                stackStore(stackTop, linkFar);
                stackStore(stackTop + WORD_STRIDE, __float_as_uint(tFar));
            }
            stackTop += ENTRY_STRIDE;
        }
        if (tNear < tmax)
            enter(linkNear);
        else
            state = POP;
    };

    // Set up the traversal of the ray (o, d, tmax) on this lane: rayAabbIntersector, wgsl:438-445 /
    // ray_intersection.cpp:92-99, and the visit of node 0.
    const auto startRay = [&]() {
        ix = __fdiv_rn(1.0f, d.x), iy = __fdiv_rn(1.0f, d.y), iz = __fdiv_rn(1.0f, d.z);
        negMask = (ix < 0.0f ? 1u : 0u) | (iy < 0.0f ? 2u : 0u) | (iz < 0.0f ? 4u : 0u);
        // +-inf inverse components (axis-parallel rays) stay on the fast path (NaN products are caught per box); NaN or zero
        // ones (NaN or infinite direction components), non-finite origins and scenes with unordered boxes do not.
        const bool exact = !sceneOrdered || !(ix == ix && iy == iy && iz == iz && ix != 0.0f && iy != 0.0f && iz != 0.0f &&
                                              isFiniteBits(o.x) && isFiniteBits(o.y) && isFiniteBits(o.z));
        if (exact) negMask |= 8u;
        rayNodes = 1, rayTris = 0, triDone = 0;
        stackTop = stackBase;
        hit.tri = RF_NO_HIT;
        const float t = slabEntry(pairs.rootBox[0], pairs.rootBox[1], pairs.rootBox[2], pairs.rootBox[3], pairs.rootBox[4], pairs.rootBox[5], negMask, o,
                                  ix, iy, iz, exact);
        if (t < tmax)
            enter(pairs.rootLink);
        else
            state = DONE;
    };

    while (true)
    {
        // ---- node rounds ----------------------------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < ROUNDS_PER_VOTE; ++k)
        {
            while (state == POP) popStep();
            if (state == EXPAND) expandStep();
        }

        const unsigned nodeMask = __ballot_sync(0xFFFFFFFFu, state == EXPAND || state == POP);
        const unsigned triMask = __ballot_sync(0xFFFFFFFFu, state == TRI);

        // ---- triangle round: once enough lanes are parked, each tests one triangle of its leaf -----------------
        if (triMask != 0u && (static_cast<std::uint32_t>(__popc(triMask)) >= tuning.triMin || nodeMask == 0u))
        {
            if (state == TRI)
            {
                const std::uint32_t tri = (cur & 0xFFFFFFu) + triDone;
                ++rayTris;
                ++triDone;
                float u, v, t;
                bool  done = false;
                if (intersectTriangle(tris, tri, o, d, tmax, u, v, t))
                {
                    hit.tri = tri, hit.u = u, hit.v = v, hit.t = t;
                    if (RF_ANY_HIT)
                        done = true; // shadowRay returns on the first accepted triangle (wgsl:340-342)
                    else
                        tmax = t;
                }
                if (done || triDone == ((cur >> 24) & 127u) + 1u)
                {
                    triDone = 0;
                    state = done ? DONE : POP;
                }
            }
            continue; // the masks are stale now; vote again after the next node rounds
        }

        // ---- hand finished rays to IO, refill idle lanes with the next work items, terminate ----------
        if ((nodeMask | triMask) == 0xFFFFFFFFu) continue;
        if (state == DONE)
        {
            if (io.finish(rayIdx, hit.tri != RF_NO_HIT, hit, rayNodes, rayTris, RF_ANY_HIT, o, d, tmax, laneAnyHit))
                startRay(); // chained ray (e.g. the closest-hit ray of a path right after its shadow ray)
            else
                state = IDLE;
        }
        const unsigned      busyMask = __ballot_sync(0xFFFFFFFFu, state != IDLE);
        const std::uint32_t idleCount = 32u - static_cast<std::uint32_t>(__popc(busyMask));
        bool                gotWork = false;
        if (!exhausted && idleCount != 0u && (idleCount >= tuning.refillMin || busyMask == 0u))
        {
            std::uint32_t       base = 0;
            const std::uint32_t granted = io.acquire(idleCount, busyMask == 0u, base, exhausted);
            gotWork = granted != 0u;
            if (state == IDLE)
            {
                const std::uint32_t rank = static_cast<std::uint32_t>(__popc(~busyMask & ((1u << laneId()) - 1u)));
                std::uint32_t       id = base + rank; // IO may translate the work-item index into its own ray id
                if (rank < granted && io.fetch(id, o, d, tmax, laneAnyHit))
                {
                    rayIdx = id;
                    startRay();
                }
            }
        }
        if (exhausted && busyMask == 0u && !gotWork) break;
    }
#undef RF_ANY_HIT
}
} // namespace rfb200

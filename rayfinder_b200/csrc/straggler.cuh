// The tail of a traversal launch, one WARP per ray.
//
// Measured on B200 (tools/trace_timeline.py, tools/lone_ray.py, tools/microbench/chase.cu): once the ray queue of
// a launch is dry, the rays still in flight are alone in their warps and walk the BVH one dependent step after the
// other — ~620 cycles per node (a 300-cycle L2 hit, 60 cycles of slab arithmetic, then stack and loop
// bookkeeping), ~0.3 us.  A grazing ray of Sponza visits 1 100 nodes and 460 leaves, so the last ten warps of a
// launch run for 0.2-0.5 ms while 148 SMs idle, nine times per frame.
//
// Here such a ray (handed over by traceRays as a StragglerRecord) gets a whole warp:
//   * WINDOW.  The nodes are stored in depth-first order (first child = idx + 1), so the next nodes a ray visits
//     are mostly the ones right behind the current one.  The warp loads the 32 nodes [cur, cur + 32) with one
//     coalesced 1 KB request (ONE L2 round trip instead of up to 32), every lane runs the tmax-independent part of
//     the slab test for its node — and, if the node is a leaf whose box the ray enters, the tmax-independent part
//     of Moeller-Trumbore for the leaf's first triangle — and parks the results in shared memory.
//   * WALK.  The traversal itself is the reference's, node by node in the reference's order, but a visit inside
//     the window is one shared-memory read and a compare against the current tmax (~70 cycles); only a jump out of
//     the window pays memory latency again.  Leaves with more than one triangle are tested one triangle per lane.
// Every fp32 operation is the one traceRays performs (same helpers), and the order of visits and of hit updates
// is unchanged, so hits, node counts and triangle counts stay bit-identical (tests: scheduling independence).
#pragma once

#include "bulk_copy.cuh"
#include "traversal.cuh"

namespace rfb200
{
constexpr int STRAGGLER_WARPS_PER_BLOCK = 4;
constexpr int STRAGGLER_WARPS_PER_SM = 32; // measured: 16 is slower, 64 no faster
constexpr int STRAGGLER_WINDOW = 32; // nodes per window = one per lane

// What a warp needs to walk one ray through 32-node windows (1152 bytes).
struct __align__(16) StragglerWindowShared
{
    uint4         node[STRAGGLER_WINDOW]; // (tmin bits, a, b, flags)
    uint4         tri[STRAGGLER_WINDOW];  // first triangle of a leaf: (valid, u bits, v bits, t bits)
    std::uint32_t stack[RF_STACK_SIZE];
};
struct __align__(16) StragglerWarpShared : StragglerWindowShared
{
    // bulk-copy mode: two staging buffers of one raw window each, filled by cp.async.bulk — [0] on demand, [1] ahead of time
    PackedNode         raw[2][STRAGGLER_WINDOW];
    unsigned long long barrier[2]; // their mbarriers
};

// How a window reaches the warp (compile time; results never depend on it):
//   STRAGGLER_DIRECT  every lane loads its node of the window with one LDG.256 (round 1)
//   STRAGGLER_BULK    the window is staged in shared memory by ONE bulk asynchronous copy of 1 KB (cp.async.bulk through the
//                     TMA unit, completion on an mbarrier), and while the warp walks a window the copy of the window it will
//                     most probably need next — the one of the newest stack entry that lies outside the current window — is
//                     already in flight into a second buffer: the L2 round trip of a window switch overlaps the walk
enum StragglerWindowMode : int
{
    STRAGGLER_DIRECT = 0,
    STRAGGLER_BULK = 1
};

constexpr std::uint32_t WIN_OK = 1u;  // tmin <= tmx && tmx > 0 (the tmax-independent part of the slab test)
constexpr std::uint32_t WIN_NAN = 2u; // a slab product was NaN: re-test with the literal form at the visit

// Walks the ray `r` (warp-uniform: every lane holds the same values; its stack entries are in sh.stack[0, r.sp)) to its end
// on the calling warp (all 32 lanes, converged) and hands the result to io.finish on lane 0.  Returns true when finish chained
// another ray (e.g. a path's closest-hit ray right after its shadow ray): `r` then holds that ray, fresh, on every lane.
// SH = StragglerWindowShared (DIRECT mode) or StragglerWarpShared.
template<int WINDOW_MODE, class SH, class IO>
__device__ __forceinline__ bool traceWarpRay(
    const PackedNode* __restrict__ nodes,
    const float4* __restrict__ tris,
    WarpRay&       r,
    SH&            sh,
    std::uint32_t& barrierParity, // bulk-copy mode: bit b = phase parity the next wait on sh.barrier[b] expects (kept by the caller across rays)
    IO&            io)
{
    const std::uint32_t lane = laneId();
    const std::uint32_t rayIdx = r.rayIdx;
    std::uint32_t       cur = r.cur, pendTri = r.pendTri, pendEnd = r.pendEnd, rayNodes = r.rayNodes, rayTris = r.rayTris;
    float               tmax = r.tmax;
    const V3            o = r.o;
    const V3            d = r.d;
    HitRecord           hit = r.hit;
    const int           state = r.state; // 1 NODE, 2 TRI, 3 DONE (traceRays)
    const bool          anyHit = r.anyHit;
    std::uint32_t       sp = r.sp;
    const float         ix = __fdiv_rn(1.0f, d.x), iy = __fdiv_rn(1.0f, d.y), iz = __fdiv_rn(1.0f, d.z);
    const std::uint32_t negMask = (ix < 0.0f ? 1u : 0u) | (iy < 0.0f ? 2u : 0u) | (iz < 0.0f ? 4u : 0u);
    __syncwarp();

#ifdef RF_TRACE_TIMELINE
    const unsigned long long tlStart = globalTimerNs();
    const std::uint32_t      tlNodes0 = rayNodes, tlTris0 = rayTris;
    std::uint32_t            tlWindows = 0;
#endif
    std::uint32_t base = 0x80000000u; // first node of the window (none yet: node indices are < 2^31, so cur - base >= 32)
    PackedNode    mine{};             // this lane's node of the window (kept for the literal re-test)
    bool          done = state == 3;

    // ---- bulk-copy mode: staging buffers and their barriers (one warp owns them; phases persist across rays) -----------
    constexpr std::uint32_t WINDOW_BYTES = STRAGGLER_WINDOW * sizeof(PackedNode);
    std::uint32_t           bar0 = 0, bar1 = 0;
    if constexpr (WINDOW_MODE == STRAGGLER_BULK) bar0 = sharedAddress(&sh.barrier[0]), bar1 = sharedAddress(&sh.barrier[1]);
    std::uint32_t           aheadBase = 0x80000000u; // window in flight into (or sitting in) raw[1]; none: 0x80000000
    const auto issueCopy = [&](const int buffer, const std::uint32_t first) {
        // every lane has finished reading the buffer (the caller synchronised the warp) before the copy overwrites it.  No
        // proxy fence here: it is needed after generic-proxy WRITES that the async proxy must see, not after reads — and on
        // sm_100a it costs an L1 invalidation (SYNCS.CCTL.IVALL), which made this kernel 7 % slower when it sat here.
        if constexpr (WINDOW_MODE == STRAGGLER_BULK)
        {
            if (lane == 0u)
            {
                const std::uint32_t bar = buffer ? bar1 : bar0;
                mbarrierArriveExpectTx(bar, WINDOW_BYTES);
                bulkCopyGlobalToShared(sharedAddress(&sh.raw[buffer][0]), nodes + first, WINDOW_BYTES, bar);
            }
        }
    };
    // Copy the window at `first` ahead of time if the second buffer is free.
    const auto copyAhead = [&](const std::uint32_t first) {
        if (WINDOW_MODE != STRAGGLER_BULK || aheadBase != 0x80000000u) return;
        __syncwarp();
        aheadBase = first;
        issueCopy(1, first);
    };

    // Window [first, first + 32): lane L takes node first + L.  (The node array is padded with 64 zeroed records,
    // which read as never-visited interior nodes, so the window may run past the last node.)
    const auto loadWindow = [&](const std::uint32_t first) {
        __syncwarp(); // nobody still reads the previous window
#ifdef RF_TRACE_TIMELINE
        ++tlWindows;
#endif
        base = first;
        if constexpr (WINDOW_MODE == STRAGGLER_BULK)
        {
            const int buffer = aheadBase == first ? 1 : 0;
            if (buffer == 0) issueCopy(0, first);
            const std::uint32_t bar = buffer ? bar1 : bar0;
            while (!mbarrierTryWait(bar, (barrierParity >> buffer) & 1u)) {}
            barrierParity ^= 1u << buffer;
            if (buffer == 1) aheadBase = 0x80000000u;
            const uint4* src = reinterpret_cast<const uint4*>(&sh.raw[buffer][lane]);
            const uint4  lo = src[0], hi = src[1];
            mine = PackedNode{__uint_as_float(lo.x), __uint_as_float(lo.y), __uint_as_float(lo.z), __uint_as_float(lo.w), __uint_as_float(hi.x), __uint_as_float(hi.y), hi.z, hi.w};
        }
        else
        {
            mine = loadNode(nodes + first + lane);
        }
        const float x0 = (mine.minX - o.x) * ix, x1 = (mine.maxX - o.x) * ix;
        const float y0 = (mine.minY - o.y) * iy, y1 = (mine.maxY - o.y) * iy;
        const float z0 = (mine.minZ - o.z) * iz, z1 = (mine.maxZ - o.z) * iz;
        const float tmin = max3Nan(minNan(x0, x1), minNan(y0, y1), minNan(z0, z1));
        const float tmx = min3Nan(maxNan(x0, x1), maxNan(y0, y1), maxNan(z0, z1));
        const bool  nan = eitherNan(tmin, tmx);
        const bool  ok = (tmin <= tmx) && (tmx > 0.0f);
        sh.node[lane] = make_uint4(__float_as_uint(tmin), mine.a, mine.b, (ok ? WIN_OK : 0u) | (nan ? WIN_NAN : 0u));
        if ((mine.b & 3u) == 3u && (ok || nan))
        {
            // first triangle of the leaf, everything but the `t < tmax` of the moment
            float      u = 0.f, v = 0.f, t = 0.f;
            const bool valid = intersectTriangle(tris, mine.a, o, d, __int_as_float(0x7F800000), u, v, t);
            sh.tri[lane] = make_uint4(valid ? 1u : 0u, __float_as_uint(u), __float_as_uint(v), __float_as_uint(t));
        }
        __syncwarp();
        // ... and while this window is walked, fetch the one the newest stack entry leads to
        if (WINDOW_MODE == STRAGGLER_BULK && sp != 0u)
        {
            const std::uint32_t top = sh.stack[sp - 1u];
            if (top - base >= static_cast<std::uint32_t>(STRAGGLER_WINDOW)) copyAhead(top);
        }
    };

    // Triangles [first, end) of a leaf, in order, against the current tmax: one triangle per lane, then the
    // sequential accept rule (closest: smallest t, the earliest on ties, because a later triangle needs t < tmax
    // strictly; any-hit: the first accepted one ends the ray).
    const auto testTriangles = [&](std::uint32_t first, const std::uint32_t end) {
        while (first < end && !done)
        {
            const std::uint32_t count = min(end - first, 32u);
            float               u = 0.f, v = 0.f, t = 0.f;
            const bool          accepted = lane < count && intersectTriangle(tris, first + lane, o, d, tmax, u, v, t);
            const unsigned      mask = __ballot_sync(0xFFFFFFFFu, accepted);
            if (anyHit)
            {
                if (mask != 0u)
                {
                    const std::uint32_t winner = static_cast<std::uint32_t>(__ffs(static_cast<int>(mask)) - 1);
                    hit.tri = first + winner;
                    hit.u = __shfl_sync(0xFFFFFFFFu, u, winner), hit.v = __shfl_sync(0xFFFFFFFFu, v, winner), hit.t = __shfl_sync(0xFFFFFFFFu, t, winner);
                    rayTris += winner + 1u;
                    done = true;
                    return;
                }
            }
            else if (mask != 0u)
            {
                // accepted t are positive floats: their bit patterns order like the values
                const std::uint32_t best = __reduce_min_sync(0xFFFFFFFFu, accepted ? __float_as_uint(t) : 0xFFFFFFFFu);
                const unsigned      ties = __ballot_sync(0xFFFFFFFFu, accepted && __float_as_uint(t) == best);
                const std::uint32_t winner = static_cast<std::uint32_t>(__ffs(static_cast<int>(ties)) - 1);
                hit.tri = first + winner;
                hit.u = __shfl_sync(0xFFFFFFFFu, u, winner), hit.v = __shfl_sync(0xFFFFFFFFu, v, winner), hit.t = __shfl_sync(0xFFFFFFFFu, t, winner);
                tmax = hit.t;
            }
            rayTris += count;
            first += count;
        }
    };

    bool needPop = false;
    if (state == 2)
    {
        // handed over in the middle of a leaf
        testTriangles(pendTri, pendEnd);
        needPop = true;
    }
    while (!done)
    {
        if (needPop)
        {
            if (sp == 0u) break;
            cur = sh.stack[--sp];
            needPop = false;
        }
        // ---- one visit: ray_intersection.cpp:156-204 -----------------------------------------------------------
        ++rayNodes;
        if (cur - base >= static_cast<std::uint32_t>(STRAGGLER_WINDOW)) loadWindow(cur);
        const std::uint32_t off = cur - base;
        const uint4         e = sh.node[off];
        bool                boxHit = (e.w & WIN_OK) != 0u && __uint_as_float(e.x) < tmax;
        if (e.w & WIN_NAN)
        {
            // rare: the owner lane re-tests its node with the literal form and the current tmax
            const bool exact = slabTestExact(mine, negMask, o, ix, iy, iz, tmax);
            boxHit = __shfl_sync(0xFFFFFFFFu, exact ? 1 : 0, off) != 0;
        }
        const std::uint32_t kind = e.z & 3u;
        if (boxHit && kind != 3u)
        {
            const bool neg = (negMask >> kind) & 1u;
            const std::uint32_t pushed = neg ? cur + 1u : e.y;
            sh.stack[sp++] = pushed; // every lane stores the same word
            cur = neg ? e.y : cur + 1u;
            // the newest entry outside the window is where the walk goes when it leaves the window by a pop
            if (WINDOW_MODE == STRAGGLER_BULK && pushed - base >= static_cast<std::uint32_t>(STRAGGLER_WINDOW)) copyAhead(pushed);
            continue;
        }
        if (boxHit)
        {
            const std::uint32_t first = e.y, end = e.y + (e.z >> 2);
            // the leaf's first triangle was tested when the window was loaded
            const uint4 p = sh.tri[off];
            const float t = __uint_as_float(p.w);
            ++rayTris;
            if (p.x != 0u && t < tmax)
            {
                hit.tri = first, hit.u = __uint_as_float(p.y), hit.v = __uint_as_float(p.z), hit.t = t;
                if (anyHit)
                    done = true;
                else
                    tmax = t;
            }
            if (!done && first + 1u < end) testTriangles(first + 1u, end);
            if (done) break;
        }
        needPop = true;
    }

    if (WINDOW_MODE == STRAGGLER_BULK && aheadBase != 0x80000000u)
    {
        // a copy the ray did not get to use is still in flight: let it land before the buffer serves the next ray
        while (!mbarrierTryWait(bar1, (barrierParity >> 1) & 1u)) {}
        barrierParity ^= 2u;
    }
    V3            o2 = o, d2 = d;
    float         tmax2 = tmax;
    bool          anyHit2 = anyHit;
    std::uint32_t id2 = rayIdx;
    int           chained = 0;
    if (lane == 0u) chained = io.finish(id2, hit.tri != RF_NO_HIT, hit, rayNodes, rayTris, anyHit, o2, d2, tmax2, anyHit2) ? 1 : 0;
#ifdef RF_TRACE_TIMELINE
    if (lane == 0u && g_timeline != nullptr)
    {
        // one record per straggler ray: tag 1, rays = nodes visited here, rounds = windows loaded, pad = triangles tested here
        const std::uint32_t at = atomicAdd(&g_timelineCount, 1u);
        if (at < g_timelineCap)
            g_timeline[at] = TimelineRecord{1ull, tlStart, tlStart, globalTimerNs(), rayNodes - tlNodes0, tlWindows, static_cast<std::uint32_t>(state), rayTris - tlTris0};
    }
#endif
    __syncwarp();
    chained = __shfl_sync(0xFFFFFFFFu, chained, 0);
    if (chained != 0)
    {
        r.rayIdx = __shfl_sync(0xFFFFFFFFu, id2, 0);
        r.o = v3(__shfl_sync(0xFFFFFFFFu, o2.x, 0), __shfl_sync(0xFFFFFFFFu, o2.y, 0), __shfl_sync(0xFFFFFFFFu, o2.z, 0));
        r.d = v3(__shfl_sync(0xFFFFFFFFu, d2.x, 0), __shfl_sync(0xFFFFFFFFu, d2.y, 0), __shfl_sync(0xFFFFFFFFu, d2.z, 0));
        r.tmax = __shfl_sync(0xFFFFFFFFu, tmax2, 0);
        r.anyHit = __shfl_sync(0xFFFFFFFFu, anyHit2 ? 1 : 0, 0) != 0;
        r.cur = 0u, r.pendTri = 0u, r.pendEnd = 0u, r.rayNodes = 0u, r.rayTris = 0u, r.sp = 0u;
        r.state = 1;
        r.hit = HitRecord{RF_NO_HIT, 0.f, 0.f, 0.f};
    }
    return chained != 0;
}

// Traces the ray of `rec` (handed over by traceRays) to its end on the calling warp.
template<int WINDOW_MODE, class IO>
__device__ __forceinline__ void traceStragglerWarp(
    const PackedNode* __restrict__ nodes,
    const float4* __restrict__ tris,
    const StragglerRecord* rec,
    StragglerWarpShared&   sh,
    std::uint32_t&         barrierParity,
    IO&                    io)
{
    // the ray, as traceRays left it (uniform: every lane reads the same words)
    const uint4 h0 = __ldcg(&rec->head[0]), h1 = __ldcg(&rec->head[1]), h2 = __ldcg(&rec->head[2]), h3 = __ldcg(&rec->head[3]),
                h4 = __ldcg(&rec->head[4]);
    if ((h0.y & 0xFFu) == 0u) return; // an empty slot (reserved by a warp that found the buffer full)
    WarpRay r;
    r.rayIdx = h0.x;
    r.cur = h0.z, r.pendTri = h0.w, r.pendEnd = h1.x, r.rayNodes = h1.y, r.rayTris = h1.z;
    r.tmax = __uint_as_float(h1.w);
    r.o = v3(__uint_as_float(h2.x), __uint_as_float(h2.y), __uint_as_float(h2.z));
    r.d = v3(__uint_as_float(h2.w), __uint_as_float(h3.x), __uint_as_float(h3.y));
    r.hit = HitRecord{h3.z, __uint_as_float(h3.w), __uint_as_float(h4.x), __uint_as_float(h4.y)};
    r.state = static_cast<int>(h0.y & 0xFFu);
    r.anyHit = (h0.y & 0x100u) != 0u;
    r.sp = h0.y >> 16;
    if (laneId() < r.sp) sh.stack[laneId()] = __ldcg(reinterpret_cast<const std::uint32_t*>(rec->stack) + laneId());
    while (traceWarpRay<WINDOW_MODE>(nodes, tris, r, sh, barrierParity, io)) {}
}
} // namespace rfb200

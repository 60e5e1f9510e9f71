// The frame as ONE persistent kernel with BLOCK-LOCAL path loops: ray generation, traversal, shading + sky and the
// compaction of live paths all run inside a single launch, and a path never leaves the block that generated it.
//
// Why: a traversal launch cannot end before its longest ray has walked its ~10^3 nodes one after the other, and at the
// share one GPU has of a 1080p frame split over 8 GPUs (260 k paths on 151 k resident lanes) every one of the staged
// pipeline's nine traversal launches is mostly that drain.  Here nothing waits for a bounce to finish: a lane that ends a
// ray takes whatever ray of whatever bounce is ready next in its block.
//
// Round 1's version of this kernel passed paths between blocks through a device-wide ring (a semaphore, head / tail
// counters and lap-tagged entries in global memory) and paid three to four dependent L2 round trips per ray for it; it
// never caught up with the staged pipeline.  This one has NO global synchronisation on the path of a ray:
//
//   block          = 7 traversal warps (traceRays with PathLoopIO) + 1 shading warp
//   path record    = 96 B in global memory (L2-resident), in a region that belongs to the block: written only by the block's
//                    shading warp, read by its traversal lanes with ONE round trip per ray (origin + the direction to trace)
//   ready ring     = shared memory, single producer (the shading warp), consumers reserve with one shared-memory CAS per
//                    warp refill; FIFO, so the paths of a block advance together and end together
//   hit ring       = shared memory, the traversal lanes' finished closest-hit rays on their way to the shading warp, which
//                    shades 32 of them at a time (one path per lane) — this is the stream compaction of live paths
//   shadow rays    = traced by the lane that then traces the path's next closest-hit ray (same origin; one 16-byte load for
//                    the new direction); the lane carries the shadow result as one bit, and the shading warp folds
//                    `radiance += contribution * visibility * invPdf` (rayColor:203) into the path's next shading step —
//                    before anything of the next bounce is added, i.e. in the reference's order.  Radiance lives in the
//                    path record and is written to the frame's buffer once, when the path ends
//   new paths      = the shading warp takes 32 pixels at a time from the frame's cursor (the only global atomic: one per 32
//                    paths) whenever 32 path slots are free, so blocks whose paths are short simply take more pixels
//   termination    = per block: cursor dry and every slot free
#pragma once

#include "kernels.cuh"

namespace rfb200
{
// Per block of BLOCK threads: up to 4 x BLOCK path slots (the launch picks how many are used), a ready ring of the same
// capacity (it can never overflow) plus one of BLOCK entries for the paths served first, a hit ring of BLOCK / 2 entries
// (every KB of shared memory is a KB less L1).
__host__ __device__ constexpr std::uint32_t loopMaxSlots(const int block) { return 4u * static_cast<std::uint32_t>(block); }
constexpr std::uint32_t LOOP_LEVELS = 16;       // bounce levels told apart by the scheduling priority (deeper ones share the last)
constexpr std::uint32_t LOOP_RECORD_VEC = 6;    // float4 per path record

// path record, float4 index
constexpr int REC_ORIGIN = 0;       // ray origin xyz, w = number of the path's next closest-hit ray (1-based, bits)
constexpr int REC_SHADOW_DIR = 1;   // direction of the shadow ray from that origin (the pixel's sun sample), w = pixel index (bits)
constexpr int REC_NEXT_DIR = 2;     // direction of the next closest-hit ray
constexpr int REC_THROUGHPUT = 3;
constexpr int REC_CONTRIBUTION = 4; // throughput * lightIntensity * reflectance of the last hit (NEE term without visibility)
constexpr int REC_RADIANCE = 5;

// ready-ring entry (16 bit): slot | flags
constexpr std::uint32_t LOOP_SLOT_MASK = 4095u;
constexpr std::uint32_t READY_SHADOW = 1u << 12;  // trace the shadow ray first
constexpr std::uint32_t READY_CLOSEST = 1u << 13; // trace the closest-hit ray (after the shadow ray, if any)
// what a traversal lane carries for its ray (traceRays' rayIdx), and word 0 of a hit-ring entry
constexpr std::uint32_t LANE_CLOSEST_FOLLOWS = 1u << 13;
constexpr std::uint32_t LANE_HAD_SHADOW = 1u << 14; // the path's shadow ray was traced on this lane just before ...
constexpr std::uint32_t LANE_SHADOW_HIT = 1u << 15; // ... and was blocked
constexpr std::uint32_t LANE_FINAL = 1u << 16;      // no closest-hit result in this entry: the path ends with its shadow ray

template<int BLOCK>
struct PathLoopShared
{
    static constexpr int WARPS = BLOCK / 32;
    static constexpr std::uint32_t MAX_SLOTS = loopMaxSlots(BLOCK), READY_CAP = MAX_SLOTS, URGENT_CAP = BLOCK, HIT_CAP = BLOCK / 2;
    static_assert(MAX_SLOTS <= LOOP_SLOT_MASK + 1u && (HIT_CAP & (HIT_CAP - 1u)) == 0u && (READY_CAP & (READY_CAP - 1u)) == 0u && (URGENT_CAP & (URGENT_CAP - 1u)) == 0u,
                  "slot bits / ring capacities");
    __host__ __device__ static constexpr std::uint32_t ringCap(const std::uint32_t ring) { return ring == 0u ? URGENT_CAP : READY_CAP; }
    uint4          hitRing[HIT_CAP];    // (lane word, tri, u, v) of finished closest-hit rays
    std::uint32_t  hitSeq[HIT_CAP];     // lap tag of the entry (position / CAP + 1), written after the entry
    std::uint16_t  urgentRing[URGENT_CAP];   // ready ring 0: paths that lag behind the others of the block, served first (when it is full they queue with the rest)
    std::uint16_t  readyRing[READY_CAP];     // ready ring 1: the rest
    std::uint16_t  freeStack[MAX_SLOTS]; // private to the shading warp
    std::uint16_t  grant[WARPS][32];         // entries a traversal warp has just reserved (acquire -> fetch)
    std::uint32_t  hitTail;                  // reserved positions (traversal lanes, atomicAdd)
    std::uint32_t  hitHead;                  // consumed positions (shading warp)
    std::uint32_t  readyHead[2];             // reserved positions (traversal warps, CAS)
    std::uint32_t  readyTail[2];             // published positions (shading warp)
    std::uint32_t  levelCount[LOOP_LEVELS];  // live paths by the number of their next closest-hit ray (shading warp only)
    std::uint32_t  tail;                     // 1: few paths are left (PathLoopIO::tailPhase)
    std::uint32_t  done;                     // 1: the block's share of the frame is complete; 2: watchdog
    std::uint32_t  blockStats[6];            // closest {rays, nodes, tris}, shadow {rays, nodes, tris}
};

__device__ __forceinline__ float4 ldcg4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ std::uint32_t volatileLoad(const std::uint32_t* p) { return *reinterpret_cast<const volatile std::uint32_t*>(p); }
__device__ __forceinline__ void volatileStore(std::uint32_t* p, const std::uint32_t v) { *reinterpret_cast<volatile std::uint32_t*>(p) = v; }

// IO of the traversal warps.  Nothing but the lane word (traceRays' rayIdx) lives in registers across the hot loop.
template<int BLOCK>
struct PathLoopIO
{
    using Shared = PathLoopShared<BLOCK>;
    static constexpr bool HANDS_OVER_STRAGGLERS = false;   // no ray leaves its warp ...
    static constexpr bool WALKS_LAST_RAY_WITH_WARP = true; // ... but at the end of the frame a warp walks its last ray with all lanes
#ifdef RF_TRACE_TIMELINE
    __device__ __forceinline__ unsigned long long timelineTag() const { return 0ull; }
#endif
    const float4* records; // this block's path records
    Shared&       sh;

    // The block's share of the frame is nearly done (set by the shading warp): warps take one ray at a time and walk it
    // with all 32 lanes.
    // (Warp-collective: the flag changes under the warp's feet, and lanes that read it at different moments must not take
    // different sides of a branch that contains warp-synchronous code.)
    __device__ __forceinline__ bool tailPhase() const { return __any_sync(0xFFFFFFFFu, volatileLoad(&sh.tail) != 0u) != 0; }

    __device__ __forceinline__ int warpId() const { return threadIdx.x >> 5; }

    // Reserve up to `want` entries of the ready ring.  The entries are read BEFORE the head moves (a reserved position may
    // be overwritten by the producer as soon as the head has passed it) and parked in sh.grant for fetch().
    __device__ __forceinline__ std::uint32_t acquire(std::uint32_t want, const bool mayWait, std::uint32_t& base, bool& exhausted) const
    {
        // at the end of the frame: one ray per warp, and only for a warp that holds none
        if (tailPhase()) want = mayWait ? 1u : 0u;
        std::uint32_t granted = 0;
        while (want != 0u)
        {
            // the ring of the lagging paths first
            std::uint32_t head = 0, n = 0, ring = 0;
            if (laneId() == 0u)
            {
                head = volatileLoad(&sh.readyHead[0]);
                n = min(want, volatileLoad(&sh.readyTail[0]) - head);
                if (n == 0u)
                {
                    ring = 1u;
                    head = volatileLoad(&sh.readyHead[1]);
                    n = min(want, volatileLoad(&sh.readyTail[1]) - head);
                }
            }
            head = __shfl_sync(0xFFFFFFFFu, head, 0);
            n = __shfl_sync(0xFFFFFFFFu, n, 0);
            ring = __shfl_sync(0xFFFFFFFFu, ring, 0);
            if (n == 0u) break;
            std::uint16_t entry = 0;
            if (laneId() < n)
            {
                const std::uint16_t* slots = ring == 0u ? sh.urgentRing : sh.readyRing;
                entry = *reinterpret_cast<const volatile std::uint16_t*>(&slots[(head + laneId()) & (Shared::ringCap(ring) - 1u)]);
            }
            std::uint32_t won = 0;
            if (laneId() == 0u) won = atomicCAS(&sh.readyHead[ring], head, head + n) == head ? 1u : 0u;
            won = __shfl_sync(0xFFFFFFFFu, won, 0);
            if (won != 0u)
            {
                sh.grant[warpId()][laneId()] = entry;
                __syncwarp();
                granted = n;
                break;
            }
        }
        if (granted == 0u)
        {
            if (__any_sync(0xFFFFFFFFu, volatileLoad(&sh.done) != 0u)) // (warp-collective, see tailPhase)
                exhausted = true;
            else if (mayWait)
                __nanosleep(200);
        }
        base = 0u; // work items are indices into sh.grant
        return granted;
    }

    __device__ __forceinline__ bool fetch(std::uint32_t& id, V3& o, V3& d, float& tmax, bool& anyHit) const
    {
        const std::uint32_t e = sh.grant[warpId()][id];
        const std::uint32_t slot = e & LOOP_SLOT_MASK;
        const float4*       rec = records + slot * LOOP_RECORD_VEC;
        anyHit = (e & READY_SHADOW) != 0u;
        const float4 oo = ldcg4(rec + REC_ORIGIN);
        const float4 dd = ldcg4(rec + (anyHit ? REC_SHADOW_DIR : REC_NEXT_DIR));
        o = v3(oo.x, oo.y, oo.z), d = v3(dd.x, dd.y, dd.z);
        tmax = 10000.0f; // T_MAX, wgsl:73
        id = slot | ((e & READY_CLOSEST) ? LANE_CLOSEST_FOLLOWS : 0u);
        return true;
    }

    __device__ __forceinline__ void pushHit(const std::uint32_t word, const HitRecord& hit) const
    {
        const std::uint32_t pos = atomicAdd(&sh.hitTail, 1u);
        std::uint32_t       spins = 0;
        while (pos - volatileLoad(&sh.hitHead) >= PathLoopShared<BLOCK>::HIT_CAP) // back-pressure (rare)
        {
            if (++spins > (1u << 24) || volatileLoad(&sh.done) == 2u)
            {
                volatileStore(&sh.done, 2u);
                return;
            }
        }
        const std::uint32_t slot = pos & (PathLoopShared<BLOCK>::HIT_CAP - 1u);
        sh.hitRing[slot] = make_uint4(word, hit.tri, __float_as_uint(hit.u), __float_as_uint(hit.v));
        __threadfence_block();
        volatileStore(&sh.hitSeq[slot], pos / PathLoopShared<BLOCK>::HIT_CAP + 1u);
    }

    __device__ __forceinline__ bool finish(
        std::uint32_t& id, const bool didHit, const HitRecord& hit, const std::uint32_t visited, const std::uint32_t tested, const bool anyHit,
        V3&, V3& d, float& tmax, bool& anyHitNext) const
    {
        std::uint32_t* st = sh.blockStats + (anyHit ? 3 : 0);
        atomicAdd(st + 0, 1u);
        atomicAdd(st + 1, visited);
        atomicAdd(st + 2, tested);
        if (anyHit)
        {
            // "Returns 1.0 if no forward intersections, 0.0 otherwise": one bit, folded into the path's radiance by the shading warp
            id |= LANE_HAD_SHADOW | (didHit ? LANE_SHADOW_HIT : 0u);
            if (id & LANE_CLOSEST_FOLLOWS)
            {
                // the same lane goes on with the path's next closest-hit ray: same origin, new direction
                const float4 dd = ldcg4(records + (id & LOOP_SLOT_MASK) * LOOP_RECORD_VEC + REC_NEXT_DIR);
                d = v3(dd.x, dd.y, dd.z);
                tmax = 10000.0f;
                anyHitNext = false;
                return true;
            }
            pushHit(id | LANE_FINAL, HitRecord{RF_NO_HIT, 0.f, 0.f, 0.f}); // last bounce: the path ends with its shadow ray
            return false;
        }
        pushHit(id, hit);
        return false;
    }
};

// A whole ray with the literal (NaN-propagating) slab test on ONE lane, its stack in `stack` (32 words of shared memory): the
// tail loop's rare path for rays the fast slab form cannot take (traceRays' traceExactRay).
template<class IO>
__device__ __forceinline__ void traceLiteralRay(const PackedNode* nodes, const float4* tris, WarpRay* ray, std::uint32_t* stack, IO* io)
{
    WarpRay&            r = *ray;
    const float         ix = __fdiv_rn(1.0f, r.d.x), iy = __fdiv_rn(1.0f, r.d.y), iz = __fdiv_rn(1.0f, r.d.z);
    const std::uint32_t negMask = (ix < 0.0f ? 1u : 0u) | (iy < 0.0f ? 2u : 0u) | (iz < 0.0f ? 4u : 0u);
    std::uint32_t       cur = 0, sp = 0, visited = 0, tested = 0;
    float               tmax = r.tmax;
    HitRecord           hit{RF_NO_HIT, 0.f, 0.f, 0.f};
    while (true)
    {
        ++visited;
        const PackedNode    nd = loadNode(nodes + cur);
        const bool          boxHit = slabTestExact(nd, negMask, r.o, ix, iy, iz, tmax);
        const std::uint32_t kind = nd.b & 3u;
        if (boxHit && kind != 3u)
        {
            const bool neg = (negMask >> kind) & 1u;
            stack[sp++] = neg ? cur + 1u : nd.a;
            cur = neg ? nd.a : cur + 1u;
            continue;
        }
        bool done = false;
        if (boxHit)
        {
            const std::uint32_t end = nd.a + (nd.b >> 2);
            for (std::uint32_t tri = nd.a; tri != end && !done; ++tri)
            {
                ++tested;
                float u, v, t;
                if (intersectTriangle(tris, tri, r.o, r.d, tmax, u, v, t))
                {
                    hit.tri = tri, hit.u = u, hit.v = v, hit.t = t;
                    if (r.anyHit)
                        done = true;
                    else
                        tmax = t;
                }
            }
        }
        if (done || sp == 0u) break;
        cur = stack[--sp];
    }
    bool anyHitNext = false;
    r.state = io->finish(r.rayIdx, hit.tri != RF_NO_HIT, hit, visited, tested, r.anyHit, r.o, r.d, r.tmax, anyHitNext) ? 1 : 0;
    r.anyHit = anyHitNext;
}

// The end of a block's share of the frame (PathLoopIO::tailPhase): the warp takes one ray at a time from the ready rings and
// walks it — and the rays chained to it — with all 32 lanes (traceWarpRay, straggler.cuh), using its own traversal-stack
// memory (>= 1152 bytes) as the walk's scratch.  `r` = the ray the warp held when it left traceRays (state 0: none).
template<int BLOCK>
__device__ __forceinline__ void pathLoopTail(const PackedNode* __restrict__ nodes, const float4* __restrict__ tris, PathLoopIO<BLOCK>& io, WarpRay& r, std::uint32_t* scratch)
{
    StragglerWindowShared& win = *reinterpret_cast<StragglerWindowShared*>(scratch);
    const std::uint32_t    lane = laneId();
    std::uint32_t          parity = 0;
    bool                   have = r.state != 0, exhausted = false;
    while (true)
    {
        if (!have)
        {
            std::uint32_t base = 0;
            if (io.acquire(1u, true, base, exhausted) == 0u)
            {
                if (exhausted) break;
                continue;
            }
            std::uint32_t id = 0;
            V3            o = v3(0.f, 0.f, 0.f), d = o;
            float         tmax = 0.f;
            bool          anyHit = false;
            if (lane == 0u) io.fetch(id, o, d, tmax, anyHit);
            r.rayIdx = __shfl_sync(0xFFFFFFFFu, id, 0);
            r.o = v3(__shfl_sync(0xFFFFFFFFu, o.x, 0), __shfl_sync(0xFFFFFFFFu, o.y, 0), __shfl_sync(0xFFFFFFFFu, o.z, 0));
            r.d = v3(__shfl_sync(0xFFFFFFFFu, d.x, 0), __shfl_sync(0xFFFFFFFFu, d.y, 0), __shfl_sync(0xFFFFFFFFu, d.z, 0));
            r.tmax = __shfl_sync(0xFFFFFFFFu, tmax, 0);
            r.anyHit = __shfl_sync(0xFFFFFFFFu, anyHit ? 1 : 0, 0) != 0;
            r.cur = 0u, r.pendTri = 0u, r.pendEnd = 0u, r.rayNodes = 0u, r.rayTris = 0u, r.sp = 0u;
            r.state = 1;
            r.hit = HitRecord{RF_NO_HIT, 0.f, 0.f, 0.f};
        }
        if (r.state == 1 && r.sp == 0u && r.cur == 0u)
        {
            // a fresh ray: one with a NaN / zero inverse direction component or a non-finite origin takes the literal slab test
            const float ix = __fdiv_rn(1.0f, r.d.x), iy = __fdiv_rn(1.0f, r.d.y), iz = __fdiv_rn(1.0f, r.d.z);
            if (!(ix == ix && iy == iy && iz == iz && ix != 0.0f && iy != 0.0f && iz != 0.0f && isFiniteBits(r.o.x) && isFiniteBits(r.o.y) && isFiniteBits(r.o.z)))
            {
                if (lane == 0u) traceLiteralRay(nodes, tris, &r, win.stack, &io);
                __syncwarp();
                have = __shfl_sync(0xFFFFFFFFu, r.state, 0) != 0;
                if (have)
                {
                    r.rayIdx = __shfl_sync(0xFFFFFFFFu, r.rayIdx, 0);
                    r.d = v3(__shfl_sync(0xFFFFFFFFu, r.d.x, 0), __shfl_sync(0xFFFFFFFFu, r.d.y, 0), __shfl_sync(0xFFFFFFFFu, r.d.z, 0));
                    r.tmax = __shfl_sync(0xFFFFFFFFu, r.tmax, 0);
                    r.anyHit = __shfl_sync(0xFFFFFFFFu, r.anyHit ? 1 : 0, 0) != 0;
                    r.cur = 0u, r.pendTri = 0u, r.pendEnd = 0u, r.rayNodes = 0u, r.rayTris = 0u, r.sp = 0u;
                    r.state = 1;
                    r.hit = HitRecord{RF_NO_HIT, 0.f, 0.f, 0.f};
                }
                continue;
            }
        }
        have = traceWarpRay<STRAGGLER_DIRECT>(nodes, tris, r, win, parity, io);
    }
}

// The shading warp: ray generation for 32 pixels at a time, rayColor:181-234 minus the traversals for 32 paths at a time,
// path slots, the ready ring, termination.
template<int BLOCK>
__device__ __forceinline__ void pathLoopShadeWarp(
    const FrameParams&   fp,
    const SceneDevice&   scene,
    const std::uint32_t* __restrict__ ownedTiles,
    float4*              records,
    const std::uint32_t  numSlots,
    float4*              radiance,
    std::uint32_t*       pixelCursor,
    unsigned long long*  stats,
    PathLoopShared<BLOCK>& sh)
{
    const V3            sunDir = v3(fp.sky.sun_direction);
    const std::uint32_t lane = laneId();
    const std::uint32_t lanesBelow = (1u << lane) - 1u;
    const std::uint32_t totalPixelSlots = fp.numOwnedTiles * TILE_PIXELS; // a multiple of 32
    // The first pixels of a block are a fixed share (groups of 32, interleaved over the blocks: block b takes groups b, b + G,
    // ...), so that a frame with few paths per block is dealt evenly; the rest is taken from the cursor as slots become free.
    const std::uint32_t staticGroups = min((totalPixelSlots / 32u) / gridDim.x, numSlots / 32u);
    const std::uint32_t dynamicBase = staticGroups * gridDim.x * 32u;
    std::uint32_t       staticTaken = 0;
    for (std::uint32_t i = lane; i < numSlots; i += 32u) sh.freeStack[i] = static_cast<std::uint16_t>(numSlots - 1u - i);
    __syncwarp();
    std::uint32_t freeTop = numSlots; // freeStack[0, freeTop) are free slots
    std::uint32_t hitHead = 0, waited = 0, generated = 0;
    std::uint32_t readyTail[2] = {0u, 0u};
    if (lane < LOOP_LEVELS) sh.levelCount[lane] = 0u;
    __syncwarp();
    bool          cursorDry = false, tailFlagged = false;
    // (tuning.tailPaths: 0 = automatic — two rays per traversal warp —, n + 1 = a threshold of n live paths, so 1 = never)
    const std::uint32_t tailThreshold = scene.tuning.tailPaths == 0u ? 2u * (BLOCK / 32 - 1) : scene.tuning.tailPaths - 1u;
    unsigned long long lastProgress = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(lastProgress));

    // Scheduling priority: longest remaining path first.  Paths whose next ray is at most one bounce ahead of the block's
    // least advanced live path go to ring 0, which the traversal warps serve first; the others wait in ring 1.  A path that
    // was held up by a long ray (~10^3 node visits against ~90 on average) then skips the queue until it has caught up,
    // instead of finishing alone long after the others.
    const auto levelOf = [](const std::uint32_t bounce) { return min(bounce, LOOP_LEVELS - 1u); };
    const auto leastAdvancedLevel = [&]() {
        const unsigned occupied = __ballot_sync(0xFFFFFFFFu, lane < LOOP_LEVELS && sh.levelCount[lane] != 0u);
        return occupied != 0u ? static_cast<std::uint32_t>(__ffs(occupied) - 1) : 0u;
    };
    // Append the paths of the lanes with `pred` to their ready rings (entries first, then the tails).
    const auto publish = [&](const bool pred, const std::uint32_t entry, const std::uint32_t level) {
        if (__ballot_sync(0xFFFFFFFFu, pred) == 0u) return;
        const std::uint32_t minLevel = leastAdvancedLevel();
        bool                urgent = level <= minLevel + 1u;
        // the urgent ring is small: what does not fit queues with the rest (harmless: when many paths are urgent, none is)
        std::uint32_t room = 0;
        if (lane == 0u) room = PathLoopShared<BLOCK>::URGENT_CAP - (readyTail[0] - volatileLoad(&sh.readyHead[0]));
        room = __shfl_sync(0xFFFFFFFFu, room, 0);
        if (static_cast<std::uint32_t>(__popc(__ballot_sync(0xFFFFFFFFu, pred && urgent))) > room) urgent = false;
        {
            const bool     mine = pred && urgent;
            const unsigned mask = __ballot_sync(0xFFFFFFFFu, mine);
            if (mine) sh.urgentRing[(readyTail[0] + static_cast<std::uint32_t>(__popc(mask & lanesBelow))) & (PathLoopShared<BLOCK>::URGENT_CAP - 1u)] = static_cast<std::uint16_t>(entry);
            readyTail[0] += static_cast<std::uint32_t>(__popc(mask));
        }
        {
            const bool     mine = pred && !urgent;
            const unsigned mask = __ballot_sync(0xFFFFFFFFu, mine);
            if (mine) sh.readyRing[(readyTail[1] + static_cast<std::uint32_t>(__popc(mask & lanesBelow))) & (PathLoopShared<BLOCK>::READY_CAP - 1u)] = static_cast<std::uint16_t>(entry);
            readyTail[1] += static_cast<std::uint32_t>(__popc(mask));
        }
        __threadfence(); // the path records (global) and the ring entries before the tails
        __syncwarp();
        if (lane < 2u) volatileStore(&sh.readyTail[lane], lane == 0u ? readyTail[0] : readyTail[1]);
    };

#ifdef RF_TRACE_TIMELINE
    const unsigned long long tlStart = globalTimerNs();
    unsigned long long       tlBusy = 0, tlMark = 0;
    std::uint32_t            tlBatches = 0, tlEntries = 0;
#endif
    while (true)
    {
#ifdef RF_TRACE_TIMELINE
        if (tlMark != 0) tlBusy += globalTimerNs() - tlMark; // the previous iteration generated or shaded
        tlMark = 0;
#endif
        // (what the other warps change is read by one lane: every lane must take the same side of the branches below)
        std::uint32_t avail = 0, waiting = 0, failed = 0;
        if (lane == 0u)
        {
            avail = volatileLoad(&sh.hitTail) - hitHead;
            waiting = (readyTail[0] - volatileLoad(&sh.readyHead[0])) + (readyTail[1] - volatileLoad(&sh.readyHead[1]));
            failed = volatileLoad(&sh.done) == 2u ? 1u : 0u;
        }
        avail = __shfl_sync(0xFFFFFFFFu, avail, 0);
        waiting = __shfl_sync(0xFFFFFFFFu, waiting, 0);
        if (__shfl_sync(0xFFFFFFFFu, failed, 0) != 0u) break;
        if (cursorDry && !tailFlagged && numSlots - freeTop <= tailThreshold)
        {
            tailFlagged = true;
            if (lane == 0u) volatileStore(&sh.tail, 1u);
        }

        // ---- new paths: 32 pixels from the frame's cursor --------------------------------------------------
        if (!cursorDry && freeTop >= 32u && (waiting < 64u || avail < 32u))
        {
#ifdef RF_TRACE_TIMELINE
            tlMark = globalTimerNs();
#endif
            std::uint32_t base = 0;
            if (staticTaken < staticGroups)
            {
                base = (blockIdx.x + staticTaken++ * gridDim.x) * 32u;
            }
            else
            {
                if (lane == 0u) base = dynamicBase + atomicAdd(pixelCursor, 32u);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base + 32u >= totalPixelSlots) cursorDry = true;
                if (base >= totalPixelSlots) continue;
            }
            std::uint32_t  px = 0, py = 0;
            const bool     valid = slotToPixel(fp, ownedTiles, base + lane, px, py);
            const unsigned mask = __ballot_sync(0xFFFFFFFFu, valid);
            std::uint32_t  slot = 0;
            if (valid)
            {
                slot = sh.freeStack[freeTop - 1u - static_cast<std::uint32_t>(__popc(mask & lanesBelow))];
                std::uint32_t idx;
                V3            origin, dir;
                primaryRay(fp, scene, px, py, idx, origin, dir);
                float4* rec = records + slot * LOOP_RECORD_VEC;
                __stcg(rec + REC_ORIGIN, make_float4(origin.x, origin.y, origin.z, __uint_as_float(1u)));
                __stcg(rec + REC_SHADOW_DIR, make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(idx)));
                __stcg(rec + REC_NEXT_DIR, make_float4(dir.x, dir.y, dir.z, 0.0f));
                __stcg(rec + REC_THROUGHPUT, make_float4(1.0f, 1.0f, 1.0f, 0.0f));
                __stcg(rec + REC_RADIANCE, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
                ++generated;
            }
            freeTop -= static_cast<std::uint32_t>(__popc(mask));
            if (lane == 0u) sh.levelCount[1] += static_cast<std::uint32_t>(__popc(mask));
            __syncwarp();
            publish(valid, slot | READY_CLOSEST, 1u);
            continue;
        }

        if (avail == 0u)
        {
            if (cursorDry && freeTop == numSlots)
            {
                if (lane == 0u) volatileStore(&sh.done, 1u);
                break;
            }
            // Watchdog: two seconds without a finished ray flag the frame as failed instead of hanging the device (a lost
            // path would be a bug; never observed).
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - lastProgress > 2000000000ull)
            {
                if (lane == 0u) volatileStore(&sh.done, 2u);
                break;
            }
            __nanosleep(200);
            continue;
        }
        if (avail < 32u && waiting >= 32u && waited < scene.tuning.shadeWait)
        {
            // the lanes have rays to go on with: let the batch fill (a dense batch costs the same as a sparse one)
            ++waited;
            __nanosleep(500);
            continue;
        }
        waited = 0u;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(lastProgress));

        // ---- shade up to 32 finished closest-hit rays ---------------------------------------------------
        const std::uint32_t n = min(avail, 32u);
#ifdef RF_TRACE_TIMELINE
        tlMark = globalTimerNs();
        ++tlBatches, tlEntries += n;
#endif
        bool                survives = false, ended = false;
        std::uint32_t       slot = 0, entry = 0, level = 0;
        if (lane < n)
        {
            const std::uint32_t pos = hitHead + lane;
            const std::uint32_t ringSlot = pos & (PathLoopShared<BLOCK>::HIT_CAP - 1u);
            std::uint32_t       spins = 0;
            while (volatileLoad(&sh.hitSeq[ringSlot]) != pos / PathLoopShared<BLOCK>::HIT_CAP + 1u) // entry being written
            {
                if (++spins > (1u << 24))
                {
                    volatileStore(&sh.done, 2u);
                    break;
                }
            }
            __threadfence_block();
            const uint4         e = sh.hitRing[ringSlot];
            const std::uint32_t word = e.x;
            slot = word & LOOP_SLOT_MASK;
            const HitRecord hit{e.y, __uint_as_float(e.z), __uint_as_float(e.w), 0.0f};
            float4*         rec = records + slot * LOOP_RECORD_VEC;
            // everything indexed by the slot is requested together: one L2 round trip
            const float4        originRay = ldcg4(rec + REC_ORIGIN);
            const std::uint32_t idx = __float_as_uint(ldcg4(rec + REC_SHADOW_DIR).w);
            const float4        thr = ldcg4(rec + REC_THROUGHPUT);
            const float4        c = ldcg4(rec + REC_CONTRIBUTION);
            float4              rad = ldcg4(rec + REC_RADIANCE);
            const float4        dir = ldcg4(rec + REC_NEXT_DIR);
            const std::uint32_t bounce = __float_as_uint(originRay.w);
            level = levelOf(bounce);
            atomicSub(&sh.levelCount[level], 1u);
            if (word & LANE_HAD_SHADOW)
            {
                // shadowRay result folded into the NEE term of the previous hit, rayColor:203
                const float vis = (word & LANE_SHADOW_HIT) ? 0.0f : 1.0f;
                rad.x += c.x * vis * fp.solarInvPdf;
                rad.y += c.y * vis * fp.solarInvPdf;
                rad.z += c.z * vis * fp.solarInvPdf;
            }
            if (word & LANE_FINAL)
            {
                ended = true;
            }
            else if (hit.tri == RF_NO_HIT)
            {
                const V3 sky = skyForMiss(fp, v3(dir.x, dir.y, dir.z), sunDir);
                rad.x += thr.x * sky.x, rad.y += thr.y * sky.y, rad.z += thr.z * sky.z;
                ended = true;
            }
            else
            {
                const SurfaceShade s = shadeSurfaceHit(fp, scene, hit, idx, v3(thr.x, thr.y, thr.z), sunDir);
                __stcg(rec + REC_ORIGIN, make_float4(s.p.x, s.p.y, s.p.z, __uint_as_float(bounce + 1u)));
                __stcg(rec + REC_SHADOW_DIR, make_float4(s.lightDir.x, s.lightDir.y, s.lightDir.z, __uint_as_float(idx)));
                __stcg(rec + REC_NEXT_DIR, make_float4(s.wi.x, s.wi.y, s.wi.z, 0.0f));
                __stcg(rec + REC_THROUGHPUT, make_float4(s.nextThroughput.x, s.nextThroughput.y, s.nextThroughput.z, 0.0f));
                __stcg(rec + REC_CONTRIBUTION, make_float4(s.contribution.x, s.contribution.y, s.contribution.z, 0.0f));
                __stcg(rec + REC_RADIANCE, rad);
                // every hit casts a shadow ray; the path goes on unless this was the last bounce (rayColor:205-207)
                entry = slot | READY_SHADOW | (bounce < fp.numBounces ? READY_CLOSEST : 0u);
                survives = true;
                level = levelOf(bounce + 1u);
                atomicAdd(&sh.levelCount[level], 1u);
            }
            if (ended) radiance[idx] = rad;
        }
        __syncwarp();
        hitHead += n;
        if (lane == 0u) volatileStore(&sh.hitHead, hitHead); // ring slots may be reused
        const unsigned endedMask = __ballot_sync(0xFFFFFFFFu, ended);
        if (ended) sh.freeStack[freeTop + static_cast<std::uint32_t>(__popc(endedMask & lanesBelow))] = static_cast<std::uint16_t>(slot);
        freeTop += static_cast<std::uint32_t>(__popc(endedMask));
        __syncwarp();
        publish(survives, entry, level);
    }
#ifdef RF_TRACE_TIMELINE
    generated = __reduce_add_sync(0xFFFFFFFFu, generated);
    if (lane == 0u && g_timeline != nullptr)
    {
        // one record per shading warp: tag 2, dry = ns spent generating / shading, rays = hit entries, rounds = batches, pad = paths
        const std::uint32_t at = atomicAdd(&g_timelineCount, 1u);
        if (at < g_timelineCap) g_timeline[at] = TimelineRecord{2ull, tlStart, tlBusy, globalTimerNs(), tlEntries, tlBatches, 0u, generated};
    }
    if (lane != 0u) generated = 0u;
#endif
    warpStatAdd(&stats[STAT_PATHS], generated);
}

template<int BLOCK, int STACK>
constexpr std::size_t megaSharedBytes() { return static_cast<std::size_t>(STACK) * BLOCK * 4u; }

template<int VARIANT, int BLOCK, int STACK = RF_STACK_SIZE>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) k_mega(
    const __grid_constant__ FrameParams fp,
    const __grid_constant__ SceneDevice scene,
    const std::uint32_t* __restrict__ ownedTiles,
    float4*             records,      // gridDim.x x slotsPerBlock x LOOP_RECORD_VEC
    const std::uint32_t slotsPerBlock,
    float4*             radiance,
    std::uint32_t*      pixelCursor,
    std::uint32_t*      failed,       // set to 1 when a block left on its watchdog
    unsigned long long* stats)
{
    using Shared = PathLoopShared<BLOCK>;
    constexpr int TRAVERSAL_WARPS = Shared::WARPS - 1;
    static_assert(TRAVERSAL_WARPS >= 1, "need at least one traversal warp and one shading warp");
    // the block's control structures are static shared memory; the traversal stacks (up to 64 KB for 512 threads, more than a
    // static allocation may have) are the kernel's dynamic shared memory (megaSharedBytes)
    __shared__ Shared       sh;
    extern __shared__ uint4 megaStackMemory[];
    std::uint32_t* const    stackMemory = reinterpret_cast<std::uint32_t*>(megaStackMemory);
    if (threadIdx.x < 6) sh.blockStats[threadIdx.x] = 0u;
    for (std::uint32_t i = threadIdx.x; i < Shared::HIT_CAP; i += BLOCK) sh.hitSeq[i] = 0u;
    if (threadIdx.x == 0) sh.hitTail = 0u, sh.hitHead = 0u, sh.readyHead[0] = sh.readyHead[1] = 0u, sh.readyTail[0] = sh.readyTail[1] = 0u, sh.done = 0u, sh.tail = 0u;
    __syncthreads();
    float4* const blockRecords = records + static_cast<std::uint64_t>(blockIdx.x) * slotsPerBlock * LOOP_RECORD_VEC;
    if (static_cast<int>(threadIdx.x >> 5) < TRAVERSAL_WARPS)
    {
        PathLoopIO<BLOCK> io{blockRecords, sh};
        static_assert(STACK * 128 >= static_cast<int>(sizeof(StragglerWindowShared)) && offsetof(StragglerWindowShared, stack) == 8 * 128, "the tail walk's scratch is the warp's own stack memory");
        WarpRay leftover;
        traceRays<2, VARIANT, BLOCK, PathLoopIO<BLOCK>, STACK, true>(scene.nodes, scene.tris, scene.ordered, scene.tuning, io, stackMemory, &leftover);
        pathLoopTail<BLOCK>(scene.nodes, scene.tris, io, leftover, stackMemory + (threadIdx.x >> 5) * (STACK * 32));
    }
    else
    {
        pathLoopShadeWarp<BLOCK>(fp, scene, ownedTiles, blockRecords, slotsPerBlock, radiance, pixelCursor, stats, sh);
    }
    __syncthreads();
    if (threadIdx.x == 0 && sh.done == 2u) atomicExch(failed, 1u);
    if (threadIdx.x < 6 && sh.blockStats[threadIdx.x] != 0u)
    {
        const int slot[6] = {STAT_CLOSEST_RAYS, STAT_CLOSEST_NODES, STAT_CLOSEST_TRIS, STAT_SHADOW_RAYS, STAT_SHADOW_NODES, STAT_SHADOW_TRIS};
        atomicAdd(&stats[slot[threadIdx.x]], static_cast<unsigned long long>(sh.blockStats[threadIdx.x]));
        if (threadIdx.x == 1 || threadIdx.x == 4) atomicAdd(&stats[STAT_RECORDS], static_cast<unsigned long long>(sh.blockStats[threadIdx.x]));
    }
}
} // namespace rfb200

// The frame as ONE persistent kernel: the stages of kernels.cuh (traversal, shading + sky, compaction of live
// paths) run inside a single launch and hand paths to each other through device queues instead of kernel
// boundaries.
//
// Why: a traversal launch cannot end before its longest ray has walked its ~10^3 nodes one after the other
// (a latency floor of ~0.25 ms per launch on B200, measured), and the staged pipeline pays that floor 9 times
// per 8-bounce frame — 2.2 ms, which is what caps multi-GPU strong scaling.  Here the floor is paid once per
// frame: while a few lanes finish long rays, the other lanes shade and trace paths of any bounce.
//
//   work item      = a path that has a ray to trace: (shadow ray of its last hit, then) its next closest-hit ray,
//                    both on the same lane, so the per-pixel order of `radiance +=` is the reference's
//                    (direct light of bounce b before anything of bounce b + 1)
//   ready ring     = device ring buffer of path ids; producers reserve with atomicAdd(tail), consumers with a
//                    counting semaphore (`avail`) + atomicAdd(head); entries carry a lap tag so a consumer can
//                    tell a published entry from a stale one without anyone resetting slots
//   shade batch    = warp specialisation: in every block all warps but one only trace; they hand the paths whose
//                    closest-hit ray they finished to the block's shading warp through a shared-memory ring, and
//                    that warp shades 32 of them at a time (dense: one path per lane), appending the survivors
//                    to the ready ring — this is the stream compaction of live paths
//   termination    = `live` counts paths that have not ended; a warp that runs dry reports the paths it ended and
//                    leaves when live == 0
//
// Path state is written by one warp and read by another without a kernel boundary in between, so every access
// to it (and to the per-frame radiance buffer) goes through L2 (`ld.global.cg`): L1 is not coherent across SMs.
#pragma once

#include "kernels.cuh"

namespace rfb200
{
struct MegaControl
{
    std::uint32_t head;  // next ring position to consume
    std::uint32_t tail;  // next ring position to produce
    int           avail; // published - reserved entries (may dip below 0 transiently)
    std::uint32_t live;  // paths that have not ended
};

constexpr std::uint32_t META_BOUNCE_MASK = 0xFFFFu; // bounce number of the path's next closest-hit ray (1-based)
constexpr std::uint32_t META_DO_CLOSEST = 1u << 16;
constexpr std::uint32_t META_DO_SHADOW = 1u << 17;
constexpr std::uint32_t RING_ID_BITS = 26; // up to 2^26 paths per sub-frame; 6 bits of lap tag
constexpr std::uint32_t RING_ID_MASK = (1u << RING_ID_BITS) - 1u;

__device__ __forceinline__ std::uint32_t ringEntry(const std::uint32_t pos, const std::uint32_t log2Cap, const std::uint32_t id)
{
    return ((((pos >> log2Cap) + 1u) & 63u) << RING_ID_BITS) | id;
}

__device__ __forceinline__ float4 ldcg4(const float4* p) { return __ldcg(p); }

// Ray generation for the persistent kernel: k_raygen + the initial ring entries and path meta.
__global__ void __launch_bounds__(BLOCK_THREADS) k_raygen_mega(
    const FrameParams fp,
    const SceneDevice scene,
    const std::uint32_t* __restrict__ ownedTiles,
    PathQueue           paths,
    std::uint32_t*      meta,
    std::uint32_t*      ready,
    const std::uint32_t log2Cap,
    std::uint32_t*      pathCount,
    float4*             radiance,
    unsigned long long* stats)
{
    const std::uint32_t total = fp.numOwnedTiles * TILE_PIXELS;
    std::uint32_t       generated = 0;
    for (std::uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < ((total + 31u) & ~31u); slot += gridDim.x * blockDim.x)
    {
        std::uint32_t       px = 0, py = 0;
        const bool          valid = slot < total && slotToPixel(fp, ownedTiles, slot, px, py);
        const std::uint32_t dst = warpAppend(pathCount, valid);
        if (!valid) continue;
        ++generated;
        std::uint32_t idx;
        V3            origin, dir;
        primaryRay(fp, scene, px, py, idx, origin, dir);
        paths.originPix[dst] = make_float4(origin.x, origin.y, origin.z, __uint_as_float(idx));
        paths.direction[dst] = make_float4(dir.x, dir.y, dir.z, 0.0f);
        paths.throughput[dst] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        meta[dst] = 1u | META_DO_CLOSEST;
        ready[dst] = ringEntry(dst, log2Cap, dst);
        radiance[idx] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    warpStatAdd(&stats[STAT_PATHS], generated);
}

__global__ void k_mega_init(MegaControl* ctl, const std::uint32_t* pathCount)
{
    const std::uint32_t n = *pathCount;
    ctl->head = 0u;
    ctl->tail = n;
    ctl->avail = static_cast<int>(n);
    ctl->live = n;
}

// ---- warp-specialised persistent kernel -------------------------------------------------------------
// A block = TRAVERSAL_WARPS warps that only trace (traceRays with MegaIO) + 1 warp that only shades.  Finished
// closest-hit rays go from the traversal lanes to the shading warp through a shared-memory ring; the shading
// warp handles 32 paths at a time (dense) and appends the survivors to the global ready ring.  Keeping the
// shading code out of the traversal warps' instruction stream keeps both under 64 registers without spills.
constexpr std::uint32_t HIT_RING_CAP = 128; // entries per block; a power of two (every KB of shared memory is a KB less L1)

template<int BLOCK>
struct MegaShared
{
    static constexpr int WARPS = BLOCK / 32;
    uint4              hitRing[HIT_RING_CAP];   // (path id, tri, u, v) of finished closest-hit rays
    std::uint32_t      hitSeq[HIT_RING_CAP];    // lap tag of the entry (position / CAP + 1), written after the entry
    std::uint32_t      ringTail;                // reserved positions (traversal lanes, atomicAdd)
    std::uint32_t      ringHead;                // consumed positions (shading warp)
    std::uint32_t      traversalAlive;          // traversal warps still running
    std::uint32_t      deadCount[WARPS];        // paths ended by the warp and not yet reported to ctl->live
    std::uint32_t      starved[WARPS];          // consecutive empty-handed waits (back-off)
    std::uint32_t      lastFailed[WARPS];       // the warp's last request got nothing: peek at the semaphore before the next one
    unsigned long long starvedSince[WARPS];     // %globaltimer (ns) of the first of them (watchdog)
    std::uint32_t      blockStats[6];           // closest {rays, nodes, tris}, shadow {rays, nodes, tris}
};

// IO of the traversal warps.  It holds no per-lane state (everything a lane needs between fetch and finish is
// re-read from the path record), so nothing but the traversal state lives in registers across the hot loop.
template<int BLOCK>
struct MegaIO
{
    using Shared = MegaShared<BLOCK>;
    static constexpr bool HANDS_OVER_STRAGGLERS = false; // rays end on the lane they started on (traversal.cuh)
#ifdef RF_TRACE_TIMELINE
    __device__ __forceinline__ unsigned long long timelineTag() const { return 0ull; }
#endif
    const FrameParams& fp;
    const SceneDevice& scene;
    const PathQueue    paths; // path state, indexed by path id
    std::uint32_t*     meta;
    float4*            radiance;
    MegaControl*       ctl;
    std::uint32_t*     ready;
    const std::uint32_t log2Cap;
    Shared&            sh;

    __device__ __forceinline__ int warpId() const { return threadIdx.x >> 5; }

    __device__ __forceinline__ std::uint32_t tryAcquire(const std::uint32_t want, std::uint32_t& base) const
    {
        std::uint32_t granted = 0, b = 0;
        // look before you leap: thousands of starving warps must not hammer the semaphore with atomics — but a warp whose
        // last request was granted goes straight to the atomic (one L2 round trip less per refill in the steady state)
        if (laneId() == 0u)
        {
            if (sh.lastFailed[warpId()] == 0u || *reinterpret_cast<volatile int*>(&ctl->avail) > 0)
            {
                const int old = atomicSub(&ctl->avail, static_cast<int>(want));
                granted = old <= 0 ? 0u : min(static_cast<std::uint32_t>(old), want);
                if (granted < want) atomicAdd(&ctl->avail, static_cast<int>(want - granted));
                if (granted != 0u) b = atomicAdd(&ctl->head, granted);
            }
            sh.lastFailed[warpId()] = granted == 0u ? 1u : 0u;
        }
        granted = __shfl_sync(0xFFFFFFFFu, granted, 0);
        base = __shfl_sync(0xFFFFFFFFu, b, 0);
        return granted;
    }

    __device__ __forceinline__ std::uint32_t acquire(const std::uint32_t want, const bool mayWait, std::uint32_t& base, bool& exhausted) const
    {
        const int           w = warpId();
        const std::uint32_t granted = tryAcquire(want, base);
        if (granted == 0u && mayWait)
        {
            // nothing to trace: report the paths this warp ended, then look at the frame
            std::uint32_t live = 0;
            if (laneId() == 0u)
            {
                const std::uint32_t ended = atomicExch(&sh.deadCount[w], 0u);
                if (ended != 0u) atomicSub(&ctl->live, ended);
                live = *reinterpret_cast<volatile std::uint32_t*>(&ctl->live);
                if (live != 0u)
                {
                    // exponential back-off (0.25 .. 8 us) keeps the polling traffic of idle warps off the L2
                    const std::uint32_t starved = sh.starved[w];
                    __nanosleep(256u << min(starved, 5u));
                    // Watchdog: a warp that has waited 2 s without the frame ending flags the frame as failed
                    // instead of hanging the device (a lost path would be a bug; never observed).
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (starved == 0u) sh.starvedSince[w] = now;
                    if (now - sh.starvedSince[w] > 2000000000ull)
                    {
                        atomicOr(&ctl->live, 0x80000000u);
                        live = 0x80000000u;
                    }
                    sh.starved[w] = starved + 1u;
                }
            }
            live = __shfl_sync(0xFFFFFFFFu, live, 0);
            exhausted = live == 0u || (live & 0x80000000u) != 0u;
        }
        else if (laneId() == 0u)
        {
            sh.starved[w] = 0u;
        }
        return granted;
    }

    __device__ __forceinline__ bool fetch(std::uint32_t& id, V3& o, V3& d, float& tmax, bool& anyHit) const
    {
        // `id` is a ring position: wait for its producer to publish it (lap tag), then take the path id
        const std::uint32_t     slot = id & ((1u << log2Cap) - 1u);
        const std::uint32_t     tag = ((id >> log2Cap) + 1u) & 63u;
        volatile std::uint32_t* entry = ready + slot;
        std::uint32_t           e = *entry;
        for (std::uint32_t spins = 0; (e >> RING_ID_BITS) != tag; ++spins)
        {
            if (spins > (1u << 22)) // watchdog, see acquire
            {
                atomicOr(&ctl->live, 0x80000000u);
                return false;
            }
            e = *entry;
        }
        id = e & RING_ID_MASK;

        const std::uint32_t m = __ldcg(meta + id);
        tmax = 10000.0f; // T_MAX, wgsl:73
        const float4 oPix = ldcg4(paths.originPix + id);
        o = v3(oPix.x, oPix.y, oPix.z);
        anyHit = (m & META_DO_SHADOW) != 0u;
        if (anyHit)
        {
            d = sunSampleDirection(fp, scene, __float_as_uint(oPix.w), v3(fp.sky.sun_direction));
        }
        else
        {
            const float4 dd = ldcg4(paths.direction + id);
            d = v3(dd.x, dd.y, dd.z);
        }
        return true;
    }

    __device__ __forceinline__ bool finish(
        const std::uint32_t id, const bool didHit, const HitRecord& hit, const std::uint32_t visited, const std::uint32_t tested, const bool anyHit,
        V3& o, V3& d, float& tmax, bool& anyHitNext) const
    {
        std::uint32_t* st = sh.blockStats + (anyHit ? 3 : 0);
        atomicAdd(st + 0, 1u);
        atomicAdd(st + 1, visited);
        atomicAdd(st + 2, tested);
        if (anyHit)
        {
            // shadowRay result folded into the NEE term, rayColor:203.  Everything that is indexed by the path id is requested
            // up front (one L2 round trip; the next direction speculatively), only the radiance word waits for the pixel index.
            const float4        oPix = ldcg4(paths.originPix + id);
            const float4        c = ldcg4(paths.contribution + id);
            const std::uint32_t m = __ldcg(meta + id);
            const float4        dd = ldcg4(paths.direction + id);
            const std::uint32_t idx = __float_as_uint(oPix.w);
            const float         vis = didHit ? 0.0f : 1.0f;
            float4              rad = ldcg4(radiance + idx);
            rad.x += c.x * vis * fp.solarInvPdf;
            rad.y += c.y * vis * fp.solarInvPdf;
            rad.z += c.z * vis * fp.solarInvPdf;
            __stcg(radiance + idx, rad);
            if (m & META_DO_CLOSEST)
            {
                // the same lane goes on with the path's next closest-hit ray
                o = v3(oPix.x, oPix.y, oPix.z);
                d = v3(dd.x, dd.y, dd.z);
                tmax = 10000.0f;
                anyHitNext = false;
                return true;
            }
            atomicAdd(&sh.deadCount[warpId()], 1u); // last bounce: the path ends with its shadow ray
            return false;
        }
        // closest-hit ray done: hand the path to the block's shading warp
        const std::uint32_t pos = atomicAdd(&sh.ringTail, 1u);
        while (pos - *reinterpret_cast<volatile std::uint32_t*>(&sh.ringHead) >= HIT_RING_CAP) {} // back-pressure (rare)
        const std::uint32_t slot = pos & (HIT_RING_CAP - 1u);
        sh.hitRing[slot] = make_uint4(id, hit.tri, __float_as_uint(hit.u), __float_as_uint(hit.v));
        __threadfence_block();
        *reinterpret_cast<volatile std::uint32_t*>(&sh.hitSeq[slot]) = pos / HIT_RING_CAP + 1u;
        return false;
    }
};

// The shading warp: rayColor:181-234 minus the traversals for 32 paths at a time, survivors appended to the
// global ready ring (stream compaction), ended paths reported to ctl->live.
template<int BLOCK>
__device__ __forceinline__ void megaShadeLoop(
    const FrameParams&  fp,
    const SceneDevice&  scene,
    const PathQueue     paths,
    std::uint32_t*      meta,
    float4*             radiance,
    MegaControl*        ctl,
    std::uint32_t*      ready,
    const std::uint32_t log2Cap,
    MegaShared<BLOCK>&  sh)
{
    const V3      sunDir = v3(fp.sky.sun_direction);
    std::uint32_t head = 0, waited = 0;
    while (true)
    {
        const std::uint32_t tail = *reinterpret_cast<volatile std::uint32_t*>(&sh.ringTail);
        const std::uint32_t alive = *reinterpret_cast<volatile std::uint32_t*>(&sh.traversalAlive);
        const std::uint32_t avail = tail - head;
        if (avail == 0u)
        {
            if (alive == 0u) break;
            __nanosleep(500);
            continue;
        }
        if (avail < 32u && alive != 0u && waited < scene.tuning.shadeWait)
        {
            // let the batch fill for up to ~8 us: a dense batch costs the same as a sparse one
            ++waited;
            __nanosleep(500);
            continue;
        }
        waited = 0u;
        const std::uint32_t n = min(avail, 32u);
        const bool          mine = laneId() < n;
        bool                survives = false, ended = false;
        std::uint32_t       id = 0;
        if (mine)
        {
            const std::uint32_t pos = head + laneId();
            const std::uint32_t slot = pos & (HIT_RING_CAP - 1u);
            while (*reinterpret_cast<volatile std::uint32_t*>(&sh.hitSeq[slot]) != pos / HIT_RING_CAP + 1u) {} // entry being written
            __threadfence_block();
            const uint4 e = sh.hitRing[slot];
            id = e.x;
            const HitRecord     hit{e.y, __uint_as_float(e.z), __uint_as_float(e.w), 0.0f};
            const std::uint32_t m = __ldcg(meta + id);
            const std::uint32_t bounce = m & META_BOUNCE_MASK;
            const float4        oPix = ldcg4(paths.originPix + id);
            const float4        thr = ldcg4(paths.throughput + id);
            const std::uint32_t idx = __float_as_uint(oPix.w);
            if (hit.tri == RF_NO_HIT)
            {
                const float4 dir = ldcg4(paths.direction + id);
                const V3     sky = skyForMiss(fp, v3(dir.x, dir.y, dir.z), sunDir);
                float4       rad = ldcg4(radiance + idx);
                rad.x += thr.x * sky.x, rad.y += thr.y * sky.y, rad.z += thr.z * sky.z;
                __stcg(radiance + idx, rad);
                ended = true;
            }
            else
            {
                const SurfaceShade s = shadeSurfaceHit(fp, scene, hit, idx, v3(thr.x, thr.y, thr.z), sunDir);
                __stcg(paths.originPix + id, make_float4(s.p.x, s.p.y, s.p.z, oPix.w));
                __stcg(paths.direction + id, make_float4(s.wi.x, s.wi.y, s.wi.z, 0.0f));
                __stcg(paths.throughput + id, make_float4(s.nextThroughput.x, s.nextThroughput.y, s.nextThroughput.z, 0.0f));
                __stcg(paths.contribution + id, make_float4(s.contribution.x, s.contribution.y, s.contribution.z, 0.0f));
                // every hit casts a shadow ray; the path goes on unless this was the last bounce (rayColor:205-207)
                __stcg(meta + id, (bounce + 1u) | META_DO_SHADOW | (bounce < fp.numBounces ? META_DO_CLOSEST : 0u));
                survives = true;
            }
        }
        __syncwarp();
        head += n;
        if (laneId() == 0u) *reinterpret_cast<volatile std::uint32_t*>(&sh.ringHead) = head; // slots may be reused
        const unsigned endedMask = __ballot_sync(0xFFFFFFFFu, ended);
        const unsigned mask = __ballot_sync(0xFFFFFFFFu, survives);
        if (mask != 0u)
        {
            const std::uint32_t count = static_cast<std::uint32_t>(__popc(mask));
            std::uint32_t       pos = 0;
            if (laneId() == 0u) pos = atomicAdd(&ctl->tail, count);
            pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
            __threadfence(); // path state before its ring entry
            if (survives)
            {
                const std::uint32_t at = pos + static_cast<std::uint32_t>(__popc(mask & ((1u << laneId()) - 1u)));
                *reinterpret_cast<volatile std::uint32_t*>(ready + (at & ((1u << log2Cap) - 1u))) = ringEntry(at, log2Cap, id);
            }
            __threadfence();
            __syncwarp();
            if (laneId() == 0u) atomicAdd(&ctl->avail, static_cast<int>(count));
        }
        if (laneId() == 0u && endedMask != 0u) atomicSub(&ctl->live, static_cast<std::uint32_t>(__popc(endedMask)));
    }
}

template<int VARIANT, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1024 / BLOCK) k_mega(
    const __grid_constant__ FrameParams fp,
    const __grid_constant__ SceneDevice scene,
    const PathQueue     paths,
    std::uint32_t*      meta,
    float4*             radiance,
    MegaControl*        ctl,
    std::uint32_t*      ready,
    const std::uint32_t log2Cap,
    unsigned long long* stats)
{
    using Shared = MegaShared<BLOCK>;
    constexpr int TRAVERSAL_WARPS = Shared::WARPS - 1;
    static_assert(TRAVERSAL_WARPS >= 1, "need at least one traversal warp and one shading warp");
    __shared__ Shared sh;
    if (threadIdx.x < 6) sh.blockStats[threadIdx.x] = 0u;
    if (threadIdx.x < Shared::WARPS) sh.deadCount[threadIdx.x] = 0u, sh.starved[threadIdx.x] = 0u, sh.lastFailed[threadIdx.x] = 0u;
    for (std::uint32_t i = threadIdx.x; i < HIT_RING_CAP; i += BLOCK) sh.hitSeq[i] = 0u;
    if (threadIdx.x == 0) sh.ringTail = 0u, sh.ringHead = 0u, sh.traversalAlive = TRAVERSAL_WARPS;
    __syncthreads();
    if (static_cast<int>(threadIdx.x >> 5) < TRAVERSAL_WARPS)
    {
        MegaIO<BLOCK> io{fp, scene, paths, meta, radiance, ctl, ready, log2Cap, sh};
        traceRays<2, VARIANT, BLOCK>(scene.nodes, scene.tris, scene.ordered, scene.tuning, io);
        // paths ended by this warp that it has not reported yet (it may have left on the watchdog)
        __syncwarp();
        if (laneId() == 0u)
        {
            const std::uint32_t ended = atomicExch(&sh.deadCount[threadIdx.x >> 5], 0u);
            if (ended != 0u) atomicSub(&ctl->live, ended);
            __threadfence_block();
            atomicSub(&sh.traversalAlive, 1u);
        }
    }
    else
    {
        megaShadeLoop<BLOCK>(fp, scene, paths, meta, radiance, ctl, ready, log2Cap, sh);
    }
    __syncthreads();
    if (threadIdx.x < 6 && sh.blockStats[threadIdx.x] != 0u)
    {
        const int slot[6] = {STAT_CLOSEST_RAYS, STAT_CLOSEST_NODES, STAT_CLOSEST_TRIS, STAT_SHADOW_RAYS, STAT_SHADOW_NODES, STAT_SHADOW_TRIS};
        atomicAdd(&stats[slot[threadIdx.x]], static_cast<unsigned long long>(sh.blockStats[threadIdx.x]));
        if (threadIdx.x == 1 || threadIdx.x == 4) atomicAdd(&stats[STAT_RECORDS], static_cast<unsigned long long>(sh.blockStats[threadIdx.x]));
    }
}
} // namespace rfb200

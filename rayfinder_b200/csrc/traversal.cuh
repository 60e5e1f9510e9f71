// BVH traversal + Moeller-Trumbore for sm_100a — the device twin of
//   rayIntersectBvh   reference_path_tracer.wgsl:371-429  / common/ray_intersection.cpp:138-213
//   shadowRay         reference_path_tracer.wgsl:323-368
//   rayIntersectAabb  wgsl:448-475 / ray_intersection.cpp:101-136
//   rayIntersectTriangle wgsl:478-521 / ray_intersection.cpp:38-90
//
// Per ray, the traversal ORDER and every fp32 operation are those of the reference (strict IEEE, no
// FMA: this file is compiled with -fmad=false), so nodesVisited is bit-exact and the closest hit — ties
// included — is the one the reference finds.  What is B200-specific is how a warp is kept busy and how
// the data is pulled in (DESIGN.md "Traversal kernel"):
//
//   * persistent warps with per-lane ray refill: a lane whose ray terminates gets the next ray of the
//     queue (one warp-aggregated atomicAdd), so lanes do not idle until the longest ray of a batch ends;
//   * node steps and triangle tests are separate phases of one loop: lanes that reached a leaf park
//     until enough lanes have triangles pending (warp vote), then one triangle round runs for all of
//     them — the frequent 32-byte node step stays convergent;
//   * node = 32 B (one sector, ONE LDG.E.256) instead of the reference's 48 B / two sectors:
//         (min.x, min.y, min.z, max.x, max.y, max.z, A, B)
//         interior: A = secondChildOffset, B = splitAxis (0..2)
//         leaf:     A = trianglesOffset,   B = (triangleCount << 2) | 3
//   * tri  = 64 B, aligned: (v0.xyz, e1.x) (e1.yz, e2.xy) (e2.z, n.xyz) (unused) with e1 = v1 - v0,
//     e2 = v2 - v0, n = normalize(cross(e1, e2)) precomputed at upload by the same fp32 operations the
//     reference performs per test; a test reads its first 36 bytes with one LDG.256 + one LDG.32;
//   * the traversal stack lives in shared memory, one column per thread (bank = lane: conflict-free for
//     any mix of stack depths), keeping the divergent push/pop traffic out of the L1 tag stage.
//
// Slab test: when no slab product (b - o) * invDir is NaN and the scene's boxes are finite and ordered
// (checked at upload), the reference's sequence of early-outs and std::max/std::min reduces exactly to
//   max3(lo) <= min3(hi) && max3(lo) < tmax && min3(hi) > 0        (proof in DESIGN.md; +-inf are fine).
// A NaN product (0 * inf) is detected per node with NaN-propagating min/max and that node is re-tested with
// the literal compare-and-select form, which reproduces the reference's NaN propagation; rays with NaN or
// infinite direction components or non-finite origins are traced entirely with the literal form.
#pragma once

#include "rf_vec.h"

#include <cstdint>
#include <cuda_runtime.h>

namespace rfb200
{
constexpr float         RF_EPSILON = 0.00001f; // wgsl:66, ray_intersection.cpp:44
constexpr int           RF_STACK_SIZE = 32;    // wgsl:327,375; ray_intersection.cpp:148
constexpr std::uint32_t RF_NO_HIT = 0xFFFFFFFFu;
constexpr int           TRACE_BLOCK_THREADS = 256;
constexpr int           TRI_STRIDE = 4; // float4 per packed triangle: (v0, e1.x) (e1.yz, e2.xy) (e2.z, n) (unused) = one 64-byte record

struct __align__(32) PackedNode
{
    float         minX, minY, minZ, maxX, maxY, maxZ;
    std::uint32_t a, b;
};
static_assert(sizeof(PackedNode) == 32, "one 32-byte sector per node");

struct __align__(16) HitRecord
{
    std::uint32_t tri; // RF_NO_HIT on miss
    float         u, v, t;
};

// A ray taken off its lane in the middle of its traversal (see "stragglers" below): everything needed to go on
// with it on another lane, bit for bit.  208 bytes, written and read as 13 x 16 B.
struct __align__(16) StragglerRecord
{
    uint4 head[5];  // (rayIdx, flags, cur, pendTri) (pendEnd, rayNodes, rayTris, tmax) (o.xyz, d.x) (d.yz, hit.tri, hit.u) (hit.v, hit.t, -, -)
    uint4 stack[8]; // the 32 stack entries, bottom first (only the first `depth` are meaningful)
};
static_assert(sizeof(StragglerRecord) == 208, "13 x 16 B");

// A ray in the form the warp-per-ray walk takes it (straggler.cuh): the same values on every lane.
struct WarpRay
{
    std::uint32_t rayIdx, cur, pendTri, pendEnd, rayNodes, rayTris, sp; // sp: stack entries (in the walk's shared-memory stack)
    int           state;                                              // 1 NODE, 2 TRI, 3 DONE
    bool          anyHit;
    float         tmax;
    V3            o, d;
    HitRecord     hit;
};

struct StragglerBuffer
{
    StragglerRecord* records;
    std::uint32_t*   count;    // records appended so far (may overshoot `capacity` transiently; readers clamp)
    std::uint32_t    capacity;
    std::uint32_t    evictMax; // a warp hands its rays over once the queue is dry and it has <= evictMax of them left (0: never)
    std::uint32_t    evictDelay; // ... and has gone through this many more loop rounds in that state: short rays end in place
};

// Work source over a plain array of `numRays` items: a device cursor advanced with one atomicAdd per warp.
struct CursorSource
{
    // Straggler policy of the IO (compile time): false = every ray ends on the lane it started on; true = once the
    // queue is dry, warps with few rays left append them to `stragglers` and exit (the IO then has that member).
    static constexpr bool HANDS_OVER_STRAGGLERS = false;
    // true = the caller goes on with the whole warp per ray once io.tailPhase() holds (see the end of traceRays' loop)
    static constexpr bool WALKS_LAST_RAY_WITH_WARP = false;
#ifdef RF_TRACE_TIMELINE
    __device__ __forceinline__ unsigned long long timelineTag() const { return reinterpret_cast<unsigned long long>(cursor); }
#endif
    std::uint32_t* cursor;
    std::uint32_t  numRays;
    __device__ __forceinline__ std::uint32_t acquire(const std::uint32_t want, bool, std::uint32_t& base, bool& exhausted) const
    {
        std::uint32_t b = 0;
        if ((threadIdx.x & 31u) == 0u) b = atomicAdd(cursor, want);
        b = __shfl_sync(0xFFFFFFFFu, b, 0);
        base = b;
        exhausted = b + want >= numRays;
        return b < numRays ? min(want, numRays - b) : 0u;
    }
};

// Scheduling knobs of the persistent loop (tunable at run time, see rf_renderer_set_tuning).
struct TraceTuning
{
    std::uint32_t triMin;    // run a triangle round once this many lanes have a triangle pending
    std::uint32_t refillMin; // refill once this many lanes are idle
    std::uint32_t shadeWait; // persistent kernel: 0.5 us naps the shading warp takes to let a batch of 32 fill (mega.cuh)
    std::uint32_t priorityMode; // persistent kernel: which paths its ready rings serve first (mega.cuh)
    std::uint32_t walkInPlace; // staged pipeline: 1 = once a launch's queue is dry, a warp left with ONE ray walks it with all 32 lanes (kernels.cuh, k_trace)
    std::uint32_t tailPaths; // persistent kernel: a block with this many live paths or fewer (and no pixels left to take) gives each of its rays a whole warp
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// One 256-bit read-only load (LDG.E.ENL2.256.CONSTANT): the whole node in a single L1 tag lookup.
__device__ __forceinline__ PackedNode loadNode(const PackedNode* p)
{
    PackedNode n;
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(n.minX), "=f"(n.minY), "=f"(n.minZ), "=f"(n.maxX), "=f"(n.maxY), "=f"(n.maxZ), "=r"(n.a), "=r"(n.b)
        : "l"(p));
    return n;
}
__device__ __forceinline__ std::uint32_t laneId() { return threadIdx.x & 31u; }

// One Moeller-Trumbore test against packed triangle `tri`.  Returns true when the reference's
// rayIntersectTriangle would (det/u/v/t windows identical), with (u, v, t) of the hit.
__device__ __forceinline__ bool intersectTriangle(
    const float4* __restrict__ tris,
    const std::uint32_t tri,
    const V3            o,
    const V3            d,
    const float         tmax,
    float&              outU,
    float&              outV,
    float&              outT)
{
    // the test needs v0, e1, e2 = the first 36 bytes of the 64-byte record: one 256-bit and one 32-bit load (two L1TEX
    // wavefronts; the 48-byte layout of round 1 took three 128-bit loads)
    float ax, ay, az, aw, bx, by, bz, bw;
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(ax), "=f"(ay), "=f"(az), "=f"(aw), "=f"(bx), "=f"(by), "=f"(bz), "=f"(bw)
        : "l"(tris + TRI_STRIDE * tri));
    const float cx = __ldg(reinterpret_cast<const float*>(tris + TRI_STRIDE * tri + 2));
    const V3    v0 = v3(ax, ay, az);
    const V3    e1 = v3(aw, bx, by);
    const V3    e2 = v3(bz, bw, cx);

    const V3    h = cross(d, e2);
    const float det = dot(e1, h);
    if (det > -RF_EPSILON && det < RF_EPSILON) return false;

    const float invDet = __fdiv_rn(1.0f, det);
    const V3    s = o - v0;
    const float u = invDet * dot(s, h);
    if (u < 0.0f || u > 1.0f) return false;

    const V3    q = cross(s, e1);
    const float v = invDet * dot(d, q);
    if (v < 0.0f || u + v > 1.0f) return false;

    const float t = invDet * dot(e2, q);
    if (t > RF_EPSILON && t < tmax)
    {
        outU = u, outV = v, outT = t;
        return true;
    }
    return false;
}

// rayIntersectAabb, literal form: bounds[dirNeg] / bounds[1 - dirNeg] selection, (b - o) * invDir, the
// two early-outs and std::max/std::min operand order (NaN-propagating) of ray_intersection.cpp:101-136.
// Out of line on purpose: only rays with a zero / non-finite direction component (or scenes with
// unordered boxes) come here, and the hot loop should not carry this code.
__device__ __noinline__ bool slabTestExact(
    const PackedNode nd, const std::uint32_t negMask, const V3 o, const float ix, const float iy, const float iz, const float rayTMax)
{
    const bool  negX = negMask & 1u, negY = negMask & 2u, negZ = negMask & 4u;
    const float loX = negX ? nd.maxX : nd.minX, hiX = negX ? nd.minX : nd.maxX;
    const float loY = negY ? nd.maxY : nd.minY, hiY = negY ? nd.minY : nd.maxY;
    const float loZ = negZ ? nd.maxZ : nd.minZ, hiZ = negZ ? nd.minZ : nd.maxZ;
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
    return boxHit && (tmin < rayTMax) && (tmx > 0.0f);
}

// NaN-propagating min / max (FMNMX.NAN, FMNMX3.NAN) and the unordered compare (FSETP.NAN).
__device__ __forceinline__ float minNan(const float a, const float b)
{
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float maxNan(const float a, const float b)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float min3Nan(const float a, const float b, const float c)
{
    float r;
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float max3Nan(const float a, const float b, const float c)
{
    float r;
    asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ bool eitherNan(const float a, const float b)
{
    int p;
    asm("{ .reg .pred q; setp.nan.f32 q, %1, %2; selp.s32 %0, 1, 0, q; }" : "=r"(p) : "f"(a), "f"(b));
    return p != 0;
}

__device__ __forceinline__ bool isFiniteBits(const float x) { return (__float_as_uint(x) & 0x7F800000u) != 0x7F800000u; }

// Shared-memory stack accessors on 32-bit shared-window addresses (one STS / LDS each).
__device__ __forceinline__ void stackStore(const std::uint32_t addr, const std::uint32_t value)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(value) : "memory");
}
__device__ __forceinline__ std::uint32_t stackLoad(const std::uint32_t addr)
{
    std::uint32_t value;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(value) : "r"(addr) : "memory");
    return value;
}

#ifdef RF_TRACE_TIMELINE
// Debug build only (python -m rayfinder_b200._build --timeline): every warp of a traversal launch appends
// (launch tag, start, queue-dry, exit, rays, node-step rounds) in %globaltimer ns; tools/trace_timeline.py reads them.
struct TimelineRecord
{
    unsigned long long tag, start, dry, exit;
    std::uint32_t      rays, rounds, sm, pad;
};
__device__ TimelineRecord* g_timeline;
__device__ std::uint32_t   g_timelineCount;
__device__ std::uint32_t   g_timelineCap;
__device__ __forceinline__ unsigned long long globalTimerNs()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Lane occupancy over time: per 16.4 us bucket of %globaltimer (absolute, 256 buckets = 4.2 ms), the number of loop rounds
// the traversal warps went through and the lanes that held a ray in them (tools/occupancy_timeline.py).
constexpr int      OCC_BUCKETS = 256;
__device__ unsigned long long g_occBusy[OCC_BUCKETS], g_occRounds[OCC_BUCKETS];
#endif

// The persistent traversal loop.  IO supplies the rays and consumes the results:
//   uint32 IO::acquire(want, mayWait, base, exhausted)   warp-uniform: reserve up to `want` work items, return
//                                            how many were granted (items base .. base+granted-1); set `exhausted`
//                                            when no item will ever come again; `mayWait` = the warp has nothing
//                                            else to do
//   bool IO::fetch(id, o, d, tmax, anyHit)   load work item `id` (in/out: IO may replace it by its own ray id;
//                                            false = skip, nothing to trace)
//   bool IO::finish(id, didHit, hit, nodesVisited, trisTested, anyHit, o, d, tmax, anyHitNext)
//                                            ray `id` is done; return true to chain another ray on this lane
//                                            (o, d, tmax, anyHitNext filled in; same id)
// anyHit = shadowRay semantics (constant rayTMax, terminate on the first accepted triangle); otherwise
// closest hit with shrinking tmax.  MODE fixes it at compile time (0: closest, 1: any-hit) or leaves it
// per ray (2: a mixed stream of closest-hit and shadow rays).  `sceneOrdered` = the upload-time check that
// every box is finite with min <= max.
// VARIANT (compile-time scheduling variant, selected per launch for tuning; results never depend on it):
//   bits 0-1: node steps per warp vote minus 1 (1..4)   bit 2: triangle round tests the whole leaf (else one
//   triangle per round)   bit 3: node step written as branches (else as selects / predication)
constexpr int TRACE_DEFAULT_VARIANT = 3;
// STACK: entries of the per-thread traversal stack.  The reference's 32 is an upper bound; a scene whose deepest
// interior node needs fewer (validateBvh) can run with a smaller shared-memory stack, which decides the SM's
// shared-memory carve-out and therefore how much L1 is left for the nodes: 4 blocks x (32 x 256 x 4 B + 1 KB) need the
// 164 KB carve-out (88 KB of L1), 31 entries fit the 132 KB one (120 KB of L1), 23 entries the 100 KB one.
// EXTERNAL_STACK: the caller supplies the STACK * BLOCK words of stack memory (a kernel whose shared memory exceeds the 48 KB
// a static allocation may have passes a piece of its dynamic shared memory); otherwise the function declares them itself.
template<int MODE, int VARIANT, int BLOCK, class IO, int STACK = 32, bool EXTERNAL_STACK = false>
__device__ __forceinline__ void traceRays(
    const PackedNode* __restrict__ nodes,
    const float4* __restrict__ tris,
    const bool          sceneOrdered,
    const TraceTuning   tuning,
    IO&                 io,
    std::uint32_t*      externalStack = nullptr,
    WarpRay*            leftover = nullptr) // IO::WALKS_LAST_RAY_WITH_WARP: the ray the warp still held when it left (state 0: none)
{
    if constexpr (IO::WALKS_LAST_RAY_WITH_WARP) leftover->state = 0;
    // Traversal stack: a [warp][entry][lane] shared array — entry k of this thread lives STACK_STRIDE * k bytes above
    // stackBase, bank = lane for every k, and the STACK x 128 bytes of a warp are contiguous.
    static_assert(STACK >= 1 && STACK <= RF_STACK_SIZE, "the reference's stack has 32 entries");
    __shared__ std::uint32_t ownStack[EXTERNAL_STACK ? 1 : STACK * BLOCK];
    std::uint32_t* const     stackMem = EXTERNAL_STACK ? externalStack : ownStack;
    constexpr std::uint32_t  STACK_STRIDE = 32u * 4u;
    const std::uint32_t      stackBase = static_cast<std::uint32_t>(__cvta_generic_to_shared(stackMem + (threadIdx.x >> 5) * (STACK * 32) + (threadIdx.x & 31u)));
    std::uint32_t            stackTop = stackBase; // address of the next free entry

    enum : int
    {
        IDLE = 0, // no ray
        NODE = 1, // next action: visit node `cur`
        TRI = 2,  // parked at a leaf: triangles [pendTri, pendEnd) to test
        DONE = 3  // traversal finished, result not yet handed to IO
    };
    constexpr int  NODE_STEPS_PER_VOTE = (VARIANT & 3) + 1;
    constexpr bool TRI_WHOLE_LEAF = (VARIANT & 4) != 0;
    constexpr bool NODE_BRANCHY = (VARIANT & 8) != 0;

    int           state = IDLE;
    std::uint32_t rayIdx = 0;
    V3            o = v3(0.f, 0.f, 0.f), d = o;
    float         ix = 0.f, iy = 0.f, iz = 0.f, tmax = 0.f;
    std::uint32_t negMask = 0; // bit a = invDir[a] < 0 (dirNeg); bit 3 stays 0
    std::uint32_t cur = 0, pendTri = 0, pendEnd = 0, rayNodes = 0, rayTris = 0;
    HitRecord     hit{RF_NO_HIT, 0.f, 0.f, 0.f};
    bool          exhausted = false;
    bool          laneAnyHit = false;
#define RF_ANY_HIT (MODE == 2 ? laneAnyHit : (MODE == 1))

    // One BVH node for a lane in NODE state (one loop iteration of ray_intersection.cpp:156-204), written
    // as selects so that it compiles to predicated straight-line code.
    const auto nodeStep = [&]() {
        ++rayNodes;
        const PackedNode nd = loadNode(nodes + cur);
        // Fast form of rayIntersectAabb (see the header comment): (b - o) * invDir for both planes of each
        // slab; the sign-selected "near" / "far" products of the reference are their min / max.  min/max are the
        // NaN-propagating variants, so a NaN product (0 * inf: an axis-parallel ray whose origin lies exactly
        // on a slab plane) surfaces in tmin/tmx, is caught by one unordered compare, and that node is re-tested
        // with the literal form.
        const float x0 = (nd.minX - o.x) * ix, x1 = (nd.maxX - o.x) * ix;
        const float y0 = (nd.minY - o.y) * iy, y1 = (nd.maxY - o.y) * iy;
        const float z0 = (nd.minZ - o.z) * iz, z1 = (nd.maxZ - o.z) * iz;
        const float tmin = max3Nan(minNan(x0, x1), minNan(y0, y1), minNan(z0, z1));
        const float tmx = min3Nan(maxNan(x0, x1), maxNan(y0, y1), maxNan(z0, z1));
        bool        boxHit = (tmin <= tmx) && (tmin < tmax) && (tmx > 0.0f);
        if (eitherNan(tmin, tmx)) boxHit = slabTestExact(nd, negMask, o, ix, iy, iz, tmax);

        const std::uint32_t kind = nd.b & 3u; // 0..2 = interior split axis, 3 = leaf
        // interior: near child first by the sign of invDir[splitAxis]; the other one is pushed
        const bool          neg = (negMask >> kind) & 1u;
        const std::uint32_t next = cur + 1u;
        if (NODE_BRANCHY)
        {
            if (boxHit && kind != 3u)
            {
                stackStore(stackTop, neg ? next : nd.a);
                stackTop += STACK_STRIDE;
                cur = neg ? nd.a : next;
            }
            else if (boxHit)
            {
                pendTri = nd.a;
                pendEnd = nd.a + (nd.b >> 2);
                state = TRI;
            }
            else if (stackTop != stackBase)
            {
                stackTop -= STACK_STRIDE;
                cur = stackLoad(stackTop);
            }
            else
            {
                state = DONE;
            }
        }
        else
        {
            const bool interior = boxHit && kind != 3u;
            const bool leaf = boxHit && kind == 3u;
            if (interior)
            {
                stackStore(stackTop, neg ? next : nd.a);
                stackTop += STACK_STRIDE;
            }
            const bool pop = !boxHit && stackTop != stackBase;
            if (pop)
            {
                stackTop -= STACK_STRIDE;
                cur = stackLoad(stackTop);
            }
            if (interior) cur = neg ? nd.a : next;
            if (leaf)
            {
                pendTri = nd.a;
                pendEnd = nd.a + (nd.b >> 2);
            }
            state = interior || pop ? NODE : (leaf ? TRI : DONE);
        }
    };

    // Whole-ray traversal with the literal (NaN-propagating) slab test; rare path, see the refill section.
    const auto traceExactRay = [&]() {
        while (true)
        {
            ++rayNodes;
            const PackedNode nd = loadNode(nodes + cur);
            const bool       boxHit = slabTestExact(nd, negMask, o, ix, iy, iz, tmax);
            const std::uint32_t kind = nd.b & 3u;
            if (boxHit && kind != 3u)
            {
                const bool neg = (negMask >> kind) & 1u;
                stackStore(stackTop, neg ? cur + 1u : nd.a);
                stackTop += STACK_STRIDE;
                cur = neg ? nd.a : cur + 1u;
                continue;
            }
            if (boxHit)
            {
                const std::uint32_t end = nd.a + (nd.b >> 2);
                for (std::uint32_t tri = nd.a; tri != end; ++tri)
                {
                    ++rayTris;
                    float u, v, t;
                    if (intersectTriangle(tris, tri, o, d, tmax, u, v, t))
                    {
                        hit.tri = tri, hit.u = u, hit.v = v, hit.t = t;
                        if (RF_ANY_HIT) return;
                        tmax = t;
                    }
                }
            }
            if (stackTop == stackBase) return;
            stackTop -= STACK_STRIDE;
            cur = stackLoad(stackTop);
        }
    };

    // Set up the traversal of the ray (o, d, tmax) on this lane: rayAabbIntersector, wgsl:438-445 /
    // ray_intersection.cpp:92-99.
    const auto startRay = [&]() {
        ix = __fdiv_rn(1.0f, d.x), iy = __fdiv_rn(1.0f, d.y), iz = __fdiv_rn(1.0f, d.z);
        negMask = (ix < 0.0f ? 1u : 0u) | (iy < 0.0f ? 2u : 0u) | (iz < 0.0f ? 4u : 0u);
        // +-inf inverse components (axis-parallel rays) stay on the fast path; NaN or zero ones (NaN or
        // infinite direction components) and non-finite origins do not.
        const bool exact = !sceneOrdered || !(ix == ix && iy == iy && iz == iz && ix != 0.0f && iy != 0.0f && iz != 0.0f &&
                                              isFiniteBits(o.x) && isFiniteBits(o.y) && isFiniteBits(o.z));
        cur = 0, rayNodes = 0, rayTris = 0;
        stackTop = stackBase;
        hit.tri = RF_NO_HIT;
        state = NODE;
        if (exact)
        {
            // Rays whose slab products can be NaN in every node (zero / non-finite direction component, non-finite
            // origin, or a scene with unordered boxes) are traced to completion right here with the literal test,
            // so the hot loop never sees them.
            traceExactRay();
            state = DONE;
        }
    };

    // ---- stragglers ------------------------------------------------------------------------------------
    // The last rays of a launch walk their ~10^3 nodes one dependent load after the other while the rest of the
    // machine idles, and each of them pins a whole block (registers, 32 KB of stack) to its SM.  With
    // IO::HANDS_OVER_STRAGGLERS a warp that is left with a few rays once the queue is dry writes their complete
    // traversal state to a device buffer and exits; a follow-up launch gives each of those rays a whole warp
    // (straggler.cuh), which walks it ~2.5x faster than a lone lane can.  The state is restored bit for bit:
    // results and counters do not change.
    const auto evictRay = [&](StragglerRecord* rec) {
        const std::uint32_t depth = (stackTop - stackBase) / STACK_STRIDE;
        const std::uint32_t flags = static_cast<std::uint32_t>(state) | (laneAnyHit ? 0x100u : 0u) | (depth << 16);
        rec->head[0] = make_uint4(rayIdx, flags, cur, pendTri);
        rec->head[1] = make_uint4(pendEnd, rayNodes, rayTris, __float_as_uint(tmax));
        rec->head[2] = make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(d.x));
        rec->head[3] = make_uint4(__float_as_uint(d.y), __float_as_uint(d.z), hit.tri, __float_as_uint(hit.u));
        rec->head[4] = make_uint4(__float_as_uint(hit.v), __float_as_uint(hit.t), 0u, 0u);
        std::uint32_t* entries = reinterpret_cast<std::uint32_t*>(rec->stack);
        for (std::uint32_t k = 0; k < depth; ++k) entries[k] = stackLoad(stackBase + k * STACK_STRIDE);
    };
    bool          mayEvict = IO::HANDS_OVER_STRAGGLERS;
    std::uint32_t evictWait = 0; // rounds spent with few rays left since the queue ran dry
#ifdef RF_TRACE_TIMELINE
    const unsigned long long tlStart = globalTimerNs();
    unsigned long long       tlDry = 0;
    std::uint32_t            tlRays = 0, tlRounds = 0, tlMaxNodes = 0;
    std::uint32_t            tlOccBucket = 0xFFFFFFFFu, tlOccBusy = 0, tlOccRounds = 0;
#endif

    while (true)
    {
#ifdef RF_TRACE_TIMELINE
        ++tlRounds;
        if (exhausted && tlDry == 0) tlDry = globalTimerNs();
        tlOccBusy += static_cast<std::uint32_t>(__popc(__ballot_sync(0xFFFFFFFFu, state != IDLE)));
        ++tlOccRounds;
        if ((tlRounds & 3u) == 0u)
        {
            const std::uint32_t bucket = static_cast<std::uint32_t>(globalTimerNs() >> 14) & (OCC_BUCKETS - 1);
            if (bucket != tlOccBucket)
            {
                if (laneId() == 0u && tlOccBucket != 0xFFFFFFFFu)
                {
                    atomicAdd(&g_occBusy[tlOccBucket], static_cast<unsigned long long>(tlOccBusy));
                    atomicAdd(&g_occRounds[tlOccBucket], static_cast<unsigned long long>(tlOccRounds));
                }
                tlOccBucket = bucket, tlOccBusy = 0u, tlOccRounds = 0u;
            }
        }
#endif
        // ---- node steps: one BVH node per lane in NODE state ------------------------------------------
#pragma unroll
        for (int k = 0; k < NODE_STEPS_PER_VOTE; ++k)
        {
            if (state == NODE) nodeStep();
        }

        const unsigned nodeMask = __ballot_sync(0xFFFFFFFFu, state == NODE);
        const unsigned triMask = __ballot_sync(0xFFFFFFFFu, state == TRI);

        // ---- triangle round: once enough lanes are parked, each tests the triangles of its leaf --------
        if (triMask != 0u && (static_cast<std::uint32_t>(__popc(triMask)) >= tuning.triMin || nodeMask == 0u))
        {
            if (state == TRI)
            {
                bool done = false;
                do
                {
                    ++rayTris;
                    float u, v, t;
                    if (intersectTriangle(tris, pendTri, o, d, tmax, u, v, t))
                    {
                        hit.tri = pendTri, hit.u = u, hit.v = v, hit.t = t;
                        if (RF_ANY_HIT)
                            done = true; // shadowRay returns on the first accepted triangle (wgsl:340-342)
                        else
                            tmax = t;
                    }
                    ++pendTri;
                } while (TRI_WHOLE_LEAF && !done && pendTri != pendEnd);
                if (!done && pendTri != pendEnd)
                {
                    // one triangle per round: stay parked for the next one
                }
                else if (!done && stackTop != stackBase)
                {
                    stackTop -= STACK_STRIDE;
                    cur = stackLoad(stackTop);
                    state = NODE;
                }
                else
                {
                    state = DONE;
                }
            }
            continue; // the masks are stale now; vote again after the next node steps
        }

        // ---- hand finished rays to IO, refill idle lanes with the next work items, terminate ----------
        if ((nodeMask | triMask) == 0xFFFFFFFFu) continue;
        if (state == DONE)
        {
#ifdef RF_TRACE_TIMELINE
            tlMaxNodes = max(tlMaxNodes, rayNodes + (rayTris << 16));
#endif
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
#ifdef RF_TRACE_TIMELINE
            tlRays += granted;
#endif
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
        if constexpr (IO::WALKS_LAST_RAY_WITH_WARP)
        {
            // The end of a persistent kernel's frame: once the IO says that only a handful of rays are left, a warp that holds
            // at most ONE ray leaves this loop, and its caller goes on ray by ray with all 32 lanes (straggler.cuh: 32-node
            // windows, ~2.5x faster than a lone lane).  The ray's state leaves through `leftover` (shuffles), its stack through
            // the first 32 words of row 8 of the warp's own stack memory (where StragglerWindowShared::stack lies).
            if (sceneOrdered && io.tailPhase(exhausted))
            {
                const unsigned walking = __ballot_sync(0xFFFFFFFFu, state == NODE || state == TRI);
                const unsigned holding = __ballot_sync(0xFFFFFFFFu, state != IDLE);
                if (holding == walking && (walking & (walking - 1u)) == 0u)
                {
                    if (walking != 0u)
                    {
                        const int src = __ffs(static_cast<int>(walking)) - 1;
                        WarpRay&  r = *leftover;
                        r.rayIdx = __shfl_sync(0xFFFFFFFFu, rayIdx, src);
                        r.cur = __shfl_sync(0xFFFFFFFFu, cur, src);
                        r.pendTri = __shfl_sync(0xFFFFFFFFu, pendTri, src), r.pendEnd = __shfl_sync(0xFFFFFFFFu, pendEnd, src);
                        r.rayNodes = __shfl_sync(0xFFFFFFFFu, rayNodes, src), r.rayTris = __shfl_sync(0xFFFFFFFFu, rayTris, src);
                        r.sp = __shfl_sync(0xFFFFFFFFu, (stackTop - stackBase) / STACK_STRIDE, src);
                        r.state = __shfl_sync(0xFFFFFFFFu, state, src);
                        r.anyHit = __shfl_sync(0xFFFFFFFFu, RF_ANY_HIT ? 1 : 0, src) != 0;
                        r.tmax = __shfl_sync(0xFFFFFFFFu, tmax, src);
                        r.o = v3(__shfl_sync(0xFFFFFFFFu, o.x, src), __shfl_sync(0xFFFFFFFFu, o.y, src), __shfl_sync(0xFFFFFFFFu, o.z, src));
                        r.d = v3(__shfl_sync(0xFFFFFFFFu, d.x, src), __shfl_sync(0xFFFFFFFFu, d.y, src), __shfl_sync(0xFFFFFFFFu, d.z, src));
                        r.hit.tri = __shfl_sync(0xFFFFFFFFu, hit.tri, src);
                        r.hit.u = __shfl_sync(0xFFFFFFFFu, hit.u, src), r.hit.v = __shfl_sync(0xFFFFFFFFu, hit.v, src), r.hit.t = __shfl_sync(0xFFFFFFFFu, hit.t, src);
                        // lane L fetches entry L of the ray's stack column, then parks it where the walk expects its stack
                        const std::uint32_t column = __shfl_sync(0xFFFFFFFFu, stackBase, src);
                        const std::uint32_t entry = laneId() < r.sp ? stackLoad(column + laneId() * STACK_STRIDE) : 0u;
                        __syncwarp();
                        stackMem[(threadIdx.x >> 5) * (STACK * 32) + 8 * 32 + laneId()] = entry;
                        __syncwarp();
                    }
                    break;
                }
            }
        }
        if constexpr (IO::HANDS_OVER_STRAGGLERS)
        {
          if (mayEvict && exhausted && !gotWork)
          {
            const std::uint32_t busy = static_cast<std::uint32_t>(__popc(busyMask));
            if (busy <= io.stragglers.evictMax && evictWait++ >= io.stragglers.evictDelay)
            {
                std::uint32_t base = 0;
                if (laneId() == 0u) base = atomicAdd(io.stragglers.count, busy);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base + busy <= io.stragglers.capacity)
                {
                    if (state != IDLE) evictRay(io.stragglers.records + base + static_cast<std::uint32_t>(__popc(busyMask & ((1u << laneId()) - 1u))));
                    break;
                }
                // Buffer full: the rays end here.  The slots this warp reserved below `capacity` are marked empty
                // (flags 0 = IDLE), so the reader does not take what an earlier launch left there for a ray.
                if (laneId() < busy && base + laneId() < io.stragglers.capacity) io.stragglers.records[base + laneId()].head[0] = make_uint4(0u, 0u, 0u, 0u);
                mayEvict = false;
            }
          }
        }
    }
#ifdef RF_TRACE_TIMELINE
    if (laneId() == 0u && tlOccBucket != 0xFFFFFFFFu)
    {
        atomicAdd(&g_occBusy[tlOccBucket], static_cast<unsigned long long>(tlOccBusy));
        atomicAdd(&g_occRounds[tlOccBucket], static_cast<unsigned long long>(tlOccRounds));
    }
    tlMaxNodes = __reduce_max_sync(0xFFFFFFFFu, tlMaxNodes & 0xFFFFu) | (__reduce_max_sync(0xFFFFFFFFu, tlMaxNodes >> 16) << 16);
    if (laneId() == 0u && g_timeline != nullptr)
    {
        const std::uint32_t at = atomicAdd(&g_timelineCount, 1u);
        std::uint32_t       sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        if (at < g_timelineCap)
            g_timeline[at] = TimelineRecord{io.timelineTag(), tlStart, tlDry, globalTimerNs(), tlRays, tlRounds, sm, tlMaxNodes};
    }
#endif
#undef RF_ANY_HIT
}

// offsetRay, wgsl:523-544 / ray_intersection.cpp:17-35 (INT_SCALE 256, FLOAT_SCALE 1/65536); the deferred
// renderer's offsetPosition (deferred_renderer_lighting_pass.wgsl:500-519) is the same function with INT_SCALE 1024
// and FLOAT_SCALE 1/16384.
__device__ __forceinline__ float offsetRayComponent(const float p, const float n, const bool deferred = false)
{
    const float ORIGIN = 1.0f / 32.0f;
    const float FLOAT_SCALE = deferred ? 1.0f / 16384.0f : 1.0f / 65536.0f;
    const float INT_SCALE = deferred ? 1024.0f : 256.0f;
    const int   off = __float2int_rz(INT_SCALE * n);
    const float po = __int_as_float(__float_as_int(p) + ((p < 0.0f) ? -off : off));
    return (fabsf(p) < ORIGIN) ? (p + FLOAT_SCALE * n) : po;
}

// Hit point of an accepted triangle: p = v0 + u*e1 + v*e2, then offsetRay(p, n) (wgsl:509-516).
__device__ __forceinline__ V3 hitPoint(const float4* __restrict__ tris, const HitRecord& hit, const bool deferred = false)
{
    const float4 a = ldg4(tris + TRI_STRIDE * hit.tri + 0);
    const float4 b = ldg4(tris + TRI_STRIDE * hit.tri + 1);
    const float4 c = ldg4(tris + TRI_STRIDE * hit.tri + 2);
    const V3     v0 = v3(a.x, a.y, a.z);
    const V3     e1 = v3(a.w, b.x, b.y);
    const V3     e2 = v3(b.z, b.w, c.x);
    const V3     n = v3(c.y, c.z, c.w);
    const V3     p = (v0 + hit.u * e1) + hit.v * e2;
    return v3(offsetRayComponent(p.x, n.x, deferred), offsetRayComponent(p.y, n.y, deferred), offsetRayComponent(p.z, n.z, deferred));
}
} // namespace rfb200

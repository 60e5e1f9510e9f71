// Wavefront stages of the B200 render path.  Each kernel restates one part of the reference's
// single fragment shader (pt/reference_path_tracer.wgsl); see DESIGN.md for the stage graph.
//
//   k_raygen      fsMain:34-55 (pixel mapping, animatedBlueNoise:603-616, generateCameraRay:237-245)
//   k_trace       rayIntersectBvh:371-429 for every live path and shadowRay:323-368 + the NEE accumulation of
//                 rayColor:203 for every hit of the previous bounce, in ONE persistent launch
//   k_shade       rayColor:181-234 minus the two traversals: hit attributes (:393-400), evalTexture
//                 (:304-307,553-565), sun sample (:288-292), Lambert scatter (:295-301), sky on miss
//                 (:213-227,248-275); appends surviving paths to the next queue (stream compaction)
//   k_accumulate  imageBuffer[idx] += rayColor(...) (fsMain:55)
//   k_display     estimator / acesFilmic / gamma (fsMain:59-63, :278-285)
#pragma once

#include "rf_internal.h"
#include "straggler.cuh"
#include "traversal.cuh"
#include "traversal_pairs.cuh"

namespace rfb200
{
constexpr int TILE = 32;           // ownership tile edge (pixels)
constexpr int TILE_PIXELS = 1024;  // TILE * TILE
constexpr int BLOCK_THREADS = 256;

// One queue of live paths (structure of float4 arrays; entry i of each array belongs to path i).
struct PathQueue
{
    float4* originPix;   // ray origin xyz, w = pixel index (bits)
    float4* direction;   // ray direction xyz (not normalised after the first bounce, as in the reference)
    float4* throughput;  // path throughput xyz
    float4* contribution; // throughput * lightIntensity * reflectance of the hit that spawned the entry
};

// Uniform block: RenderParamsLayout (reference_path_tracer.cpp:35-120) + AlignedSkyState + constants.
struct FrameParams
{
    std::uint32_t width, height;
    std::uint32_t frameCount;
    std::uint32_t sampleIndex; // frameCount % numSamplesPerPixel
    std::uint32_t numBounces;
    std::uint32_t numOwnedTiles;
    std::uint32_t tilesX;
    std::uint32_t numTextures;
    std::uint64_t numTexels;
    rf_camera     camera;
    rf_sky_state  sky;
    float         solarCosThetaMax;
    float         solarInvPdf;
    std::uint32_t deferred; // 1: the deferred renderer's lighting pass (deferred.cuh): its offsetPosition constants, the solar
                            // disk in the sky, and `radiance += throughput * (lightIntensity * reflectance * vis * invPdf)`
};

struct SceneDevice
{
    const PackedNode*    nodes;      // 32 B per node
    const float4*        tris;       // TRI_STRIDE x float4 per triangle
    const float4*        vattr;      // 5 x float4 per triangle (VertexAttributes as-is)
    const uint4*         texDesc;    // (width, height, offset, 0)
    const std::uint32_t* texels;     // BGRA8
    const uchar2*        blueNoise;  // 128 x 128
    const SampleLutRow*  lut;        // one row per sample index
    const float*         srgbLut;    // 256
    bool                 ordered;    // every node box finite with min <= max (enables the NaN-free slab test)
    TraceTuning          tuning;
    PairSceneDevice      pairs;      // child-pair records (records == nullptr: the scene has none, see pair_records.h)
};

// Device counters (u32 array, zeroed once at the start of every frame):
//   [b]                       b = 0..numBounces: entries of the queue produced for bounce b + 1
//                             ([0] = primary rays from k_raygen, [b] = hits appended by k_shade of bounce b)
//   [numBounces + 1 + k]      work-fetch cursor of the k-th traversal launch of the frame
//   [2 numBounces + 2 + k]    straggler records appended by the k-th traversal launch
//   [3 numBounces + 3 + k]    work-fetch cursor of its follow-up launch (k_trace_stragglers)
__host__ __device__ inline std::uint32_t counterSlots(std::uint32_t numBounces) { return 4u * numBounces + 4u; }

enum StatSlot
{
    STAT_PATHS = 0,
    STAT_CLOSEST_RAYS,
    STAT_SHADOW_RAYS,
    STAT_CLOSEST_NODES,
    STAT_CLOSEST_TRIS,
    STAT_SHADOW_NODES,
    STAT_SHADOW_TRIS,
    STAT_RECORDS, // BVH records loaded by the traversal kernels (pair records, or nodes with the per-node kernel)
    STAT_FAILED,  // low word set to 1 by the persistent kernel's watchdog (mega.cuh)
    STAT_COUNT
};

// Warp-aggregated append: returns the slot for this lane if `pred`, using one atomic per warp.
__device__ __forceinline__ std::uint32_t warpAppend(std::uint32_t* counter, const bool pred)
{
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, pred);
    if (mask == 0u) return 0u;
    const int     leader = __ffs(mask) - 1;
    std::uint32_t base = 0;
    if (static_cast<int>(laneId()) == leader) base = atomicAdd(counter, static_cast<std::uint32_t>(__popc(mask)));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + static_cast<std::uint32_t>(__popc(mask & ((1u << laneId()) - 1u)));
}

__device__ __forceinline__ void warpStatAdd(unsigned long long* stat, std::uint32_t value)
{
    const std::uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, value);
    if (laneId() == 0 && sum != 0u) atomicAdd(stat, static_cast<unsigned long long>(sum));
}

// WGSL fract(x) = x - floor(x).
__device__ __forceinline__ float wgslFract(const float x) { return x - floorf(x); }

// pixarOnb (wgsl:310-319) applied to a local vector: mat3x3(u, v, n) * l = l.x*u + l.y*v + l.z*n.
__device__ __forceinline__ V3 onbTransform(const V3 n, const V3 l)
{
    const float s = (n.z >= 0.0f) ? 1.0f : -1.0f;
    const float a = __fdiv_rn(-1.0f, s + n.z);
    const float b = n.x * n.y * a;
    const V3    u = v3(1.0f + s * n.x * n.x * a, s * b, -s * n.x);
    const V3    v = v3(b, s + n.y * n.y * a, -n.y);
    return (l.x * u + l.y * v) + l.z * n;
}

// Path slot -> pixel.  Paths are enumerated tile by tile (32x32 ownership tiles); inside a tile each
// warp covers an 8x4 pixel block so that the 32 primary rays of a warp are spatially coherent.
__device__ __forceinline__ bool slotToPixel(
    const FrameParams& fp,
    const std::uint32_t* __restrict__ ownedTiles,
    const std::uint32_t slot,
    std::uint32_t&      px,
    std::uint32_t&      py)
{
    const std::uint32_t tile = ownedTiles[slot >> 10];
    const std::uint32_t within = slot & 1023u;
    const std::uint32_t warp = within >> 5, lane = within & 31u;
    px = (tile % fp.tilesX) * TILE + (warp & 3u) * 8u + (lane & 7u);
    py = (tile / fp.tilesX) * TILE + (warp >> 2) * 4u + (lane >> 3);
    return px < fp.width && py < fp.height;
}

// ---------------------------------------------------------------------------------------------
// The primary ray of fragment (px, py): vsMain/fsMain:10-17,34-54, animatedBlueNoise:603-616 (tabulated per
// sample index), generateCameraRay:237-245 with the thin lens (pointInUnitDisk:596-600).
__device__ __forceinline__ void primaryRay(
    const FrameParams& fp, const SceneDevice& scene, const std::uint32_t px, const std::uint32_t py, std::uint32_t& idx, V3& origin, V3& dir)
{
    // fragment centre -> texCoord -> coord
    const float u = __fdiv_rn(static_cast<float>(px) + 0.5f, static_cast<float>(fp.width));
    const float v = __fdiv_rn(static_cast<float>(py) + 0.5f, static_cast<float>(fp.height));
    const std::uint32_t cx = static_cast<std::uint32_t>(u * static_cast<float>(fp.width));
    const std::uint32_t cy = static_cast<std::uint32_t>(v * static_cast<float>(fp.height));
    idx = cy * fp.width + cx;

    const uchar2        bn = scene.blueNoise[(cy % BLUE_NOISE_HEIGHT) * BLUE_NOISE_WIDTH + (cx % BLUE_NOISE_WIDTH)];
    const SampleLutRow& lut = scene.lut[fp.sampleIndex];
    const float         ux = lut.ux[bn.x], uy = lut.uy[bn.y];

    // jitter = blueNoise / vec2f(dimensions); generateCameraRay(bn, camera, u + j.x, (1 - v) + j.y)
    const float su = u + __fdiv_rn(ux, static_cast<float>(fp.width));
    const float sv = (1.0f - v) + __fdiv_rn(uy, static_cast<float>(fp.height));

    const float r = __fsqrt_rn(ux);
    const float lensX = fp.camera.lens_radius * (r * lut.cosPhi[bn.y]);
    const float lensY = fp.camera.lens_radius * (r * lut.sinPhi[bn.y]);
    const V3    lensOffset = lensX * v3(fp.camera.right) + lensY * v3(fp.camera.up);
    origin = v3(fp.camera.origin) + lensOffset;
    dir = normalize(((v3(fp.camera.lower_left_corner) + su * v3(fp.camera.horizontal)) + sv * v3(fp.camera.vertical)) - origin);
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_raygen(
    const FrameParams fp,
    const SceneDevice scene,
    const std::uint32_t* __restrict__ ownedTiles,
    PathQueue           out,
    std::uint32_t*      outCount,
    float4*             radiance,
    unsigned long long* stats)
{
    const std::uint32_t total = fp.numOwnedTiles * TILE_PIXELS;
    std::uint32_t       generated = 0;
    for (std::uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < ((total + 31u) & ~31u);
         slot += gridDim.x * blockDim.x)
    {
        std::uint32_t px = 0, py = 0;
        const bool    valid = slot < total && slotToPixel(fp, ownedTiles, slot, px, py);
        const std::uint32_t dst = warpAppend(outCount, valid);
        if (!valid) continue;
        ++generated;
        std::uint32_t idx;
        V3            origin, dir;
        primaryRay(fp, scene, px, py, idx, origin, dir);
        out.originPix[dst] = make_float4(origin.x, origin.y, origin.z, __uint_as_float(idx));
        out.direction[dst] = make_float4(dir.x, dir.y, dir.z, 0.0f);
        out.throughput[dst] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
        radiance[idx] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    warpStatAdd(&stats[STAT_PATHS], generated);
}

// skyRadiance, wgsl:248-275 (miss path; no solar disk term).
__device__ __forceinline__ float skyRadianceChannel(const rf_sky_state& sky, const float theta, const float gamma, const int ch)
{
    const float  r = sky.sky_radiances[ch];
    const float* p = sky.params + 9 * ch;
    const float  cosGamma = cosf(gamma);
    const float  cosGamma2 = cosGamma * cosGamma;
    const float  cosTheta = fabsf(cosf(theta));
    const float  expM = expf(p[4] * gamma);
    const float  rayM = cosGamma2;
    const float  mieMLhs = 1.0f + cosGamma2;
    const float  mieMRhs = powf(1.0f + p[8] * p[8] - 2.0f * p[8] * cosGamma, 1.5f);
    const float  mieM = __fdiv_rn(mieMLhs, mieMRhs);
    const float  zenith = __fsqrt_rn(cosTheta);
    const float  radianceLhs = 1.0f + p[0] * expf(__fdiv_rn(p[1], cosTheta + 0.01f));
    const float  radianceRhs = p[2] + p[3] * expM + p[5] * rayM + p[6] * mieM + p[7] * zenith;
    return r * (radianceLhs * radianceRhs);
}
// The deferred renderer's skyRadiance adds the solar disk (deferred_renderer_lighting_pass.wgsl:229-235).
__device__ __forceinline__ float solarDiskRadiance(const rf_sky_state& sky, const float gamma, const int ch)
{
    const float TERRESTRIAL_SOLAR_RADIUS = 0.255f * (3.1415927f / 180.0f);
    return __fdiv_rn(gamma, TERRESTRIAL_SOLAR_RADIUS) <= 1.0f ? sky.solar_radiances[ch] : 0.0f;
}

// rayColor miss branch, wgsl:212-229: sky radiance along the (possibly unnormalised) direction v.
__device__ __forceinline__ V3 skyForMiss(const FrameParams& fp, const V3 v, const V3 sunDir)
{
    const float theta = acosf(v.y);
    float       cosSun = dot(v, sunDir);
    cosSun = fminf(fmaxf(cosSun, -1.0f), 1.0f);
    const float gamma = acosf(cosSun);
    V3          sky = v3(skyRadianceChannel(fp.sky, theta, gamma, 0), skyRadianceChannel(fp.sky, theta, gamma, 1), skyRadianceChannel(fp.sky, theta, gamma, 2));
    if (fp.deferred) sky = sky + v3(solarDiskRadiance(fp.sky, gamma, 0), solarDiskRadiance(fp.sky, gamma, 1), solarDiskRadiance(fp.sky, gamma, 2));
    return sky;
}

// rayColor hit branch, wgsl:191-211, for the final (closest) accepted triangle of a path at pixel `idx`:
// hit point + attributes (wgsl:393-400), albedo (:553-565), sun sample (:288-292), NEE term without
// visibility (:196-203), Lambert scatter (:295-301).
struct SurfaceShade
{
    V3 p, wi, nextThroughput, contribution;
    V3 lightDir; // the pixel's sun sample (= sunSampleDirection): direction of the shadow ray from p
};
__device__ __forceinline__ SurfaceShade shadeSurfaceHit(
    const FrameParams& fp, const SceneDevice& scene, const HitRecord& hit, const std::uint32_t idx, const V3 throughput, const V3 sunDir)
{
    SurfaceShade out;
    // Intersection of the final (closest) accepted triangle, wgsl:393-400.
    out.p = hitPoint(scene.tris, hit, fp.deferred != 0u);
    const float  b0 = 1.0f - hit.u - hit.v, b1 = hit.u, b2 = hit.v;
    const float4 a0 = ldg4(scene.vattr + 5 * hit.tri + 0);
    const float4 a1 = ldg4(scene.vattr + 5 * hit.tri + 1);
    const float4 a2 = ldg4(scene.vattr + 5 * hit.tri + 2);
    const float4 a3 = ldg4(scene.vattr + 5 * hit.tri + 3);
    const float4 a4 = ldg4(scene.vattr + 5 * hit.tri + 4);
    const V3     nrm = (b0 * v3(a0.x, a0.y, a0.z) + b1 * v3(a1.x, a1.y, a1.z)) + b2 * v3(a2.x, a2.y, a2.z);
    const float  tu = (b0 * a3.x + b1 * a3.z) + b2 * a4.x;
    const float  tv = (b0 * a3.y + b1 * a3.w) + b2 * a4.y;
    std::uint32_t texIdx = __float_as_uint(a4.z);
    texIdx = min(texIdx, fp.numTextures - 1u); // robust buffer access

    // textureLookup, wgsl:553-565.
    const uint4         desc = scene.texDesc[texIdx];
    const float         fu = wgslFract(tu), fv = wgslFract(tv);
    const std::uint32_t tj = static_cast<std::uint32_t>(fu * static_cast<float>(desc.x));
    const std::uint32_t ti = static_cast<std::uint32_t>(fv * static_cast<float>(desc.y));
    std::uint64_t       texel = static_cast<std::uint64_t>(desc.z) + static_cast<std::uint64_t>(ti * desc.x + tj);
    texel = texel < fp.numTexels ? texel : fp.numTexels - 1u; // robust buffer access clamps
    const std::uint32_t bgra = __ldg(scene.texels + texel);
    const V3            albedo = v3(scene.srgbLut[(bgra >> 16) & 0xFFu], scene.srgbLut[(bgra >> 8) & 0xFFu], scene.srgbLut[bgra & 0xFFu]);

    // Per-pixel sample (the same vec2 for every decision of the path, wgsl:194,209).
    const std::uint32_t cx = idx % fp.width, cy = idx / fp.width;
    const uchar2        bn = scene.blueNoise[(cy % BLUE_NOISE_HEIGHT) * BLUE_NOISE_WIDTH + (cx % BLUE_NOISE_WIDTH)];
    const SampleLutRow& lut = scene.lut[fp.sampleIndex];
    const float         ux = lut.ux[bn.x];
    const float         cosPhi = lut.cosPhi[bn.y], sinPhi = lut.sinPhi[bn.y];

    // sampleSolarDiskDirection -> directionInCone, wgsl:288-292,569-579.
    const float cosTheta = 1.0f - ux * (1.0f - fp.solarCosThetaMax);
    const float sinTheta = __fsqrt_rn(1.0f - cosTheta * cosTheta);
    const V3    lightDir = onbTransform(sunDir, v3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
    out.lightDir = lightDir;

    // rayColor hit branch, wgsl:191-203.
    const float FRAC_1_PI = 0.31830987f;
    const V3    lightIntensity = v3(fp.sky.solar_radiances[0], fp.sky.solar_radiances[1], fp.sky.solar_radiances[2]);
    const V3    brdf = albedo * FRAC_1_PI;
    const V3    reflectance = brdf * dot(nrm, lightDir);
    out.contribution = (throughput * lightIntensity) * reflectance;
    if (fp.deferred)
    {
        // lightSample (deferred_renderer_lighting_pass.wgsl:188-200) keeps the throughput outside: the traversal launch adds
        // throughput * (lightIntensity * reflectance * visibility * SOLAR_INV_PDF); bounce 1 is the last one (NUM_BOUNCES = 2)
        out.contribution = lightIntensity * reflectance;
        out.wi = v3(0.f, 0.f, 0.f);
        out.nextThroughput = throughput;
        return out;
    }

    // evalImplicitLambertian -> directionInCosineWeightedHemisphere, wgsl:295-301,583-592.
    const float hemiSin = __fsqrt_rn(1.0f - ux);
    out.wi = onbTransform(nrm, v3(cosPhi * hemiSin, sinPhi * hemiSin, __fsqrt_rn(ux)));
    out.nextThroughput = throughput * albedo;
    return out;
}

// The sun sample of pixel `idx` (sampleSolarDiskDirection, wgsl:288-292): direction of its shadow rays.
__device__ __forceinline__ V3 sunSampleDirection(const FrameParams& fp, const SceneDevice& scene, const std::uint32_t idx, const V3 sunDir)
{
    const std::uint32_t cx = idx % fp.width, cy = idx / fp.width;
    const uchar2        bn = scene.blueNoise[(cy % BLUE_NOISE_HEIGHT) * BLUE_NOISE_WIDTH + (cx % BLUE_NOISE_WIDTH)];
    const SampleLutRow& lut = scene.lut[fp.sampleIndex];
    const float         ux = lut.ux[bn.x];
    const float         cosTheta = 1.0f - ux * (1.0f - fp.solarCosThetaMax);
    const float         sinTheta = __fsqrt_rn(1.0f - cosTheta * cosTheta);
    return onbTransform(sunDir, v3(lut.cosPhi[bn.y] * sinTheta, lut.sinPhi[bn.y] * sinTheta, cosTheta));
}

// ---------------------------------------------------------------------------------------------
// Shading of queue `in` after k_closest.  Misses add the sky and retire; hits fetch attributes and
// albedo, compute the NEE contribution and the scattered ray, and are appended to `out`.
__global__ void __launch_bounds__(BLOCK_THREADS) k_shade(
    const FrameParams    fp,
    const SceneDevice    scene,
    const PathQueue      in,
    const std::uint32_t* __restrict__ inCount,
    const HitRecord* __restrict__ hits,
    PathQueue            out,
    std::uint32_t*       outCount,
    float4*              radiance)
{
    const std::uint32_t n = *inCount;
    const std::uint32_t nPadded = (n + 31u) & ~31u;
    const V3            sunDir = v3(fp.sky.sun_direction);
    for (std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nPadded; i += gridDim.x * blockDim.x)
    {
        const bool live = i < n;
        bool       isHit = false;
        HitRecord  hit{};
        float4     oPix = make_float4(0.f, 0.f, 0.f, 0.f), thr = oPix, dir = oPix;
        if (live)
        {
            hit = hits[i];
            oPix = in.originPix[i];
            thr = in.throughput[i];
            isHit = hit.tri != RF_NO_HIT;
            if (!isHit)
            {
                dir = in.direction[i];
            }
        }
        const std::uint32_t idx = __float_as_uint(oPix.w);

        if (live && !isHit)
        {
            const V3 sky = skyForMiss(fp, v3(dir.x, dir.y, dir.z), sunDir);
            float4   rad = radiance[idx];
            rad.x += thr.x * sky.x, rad.y += thr.y * sky.y, rad.z += thr.z * sky.z;
            radiance[idx] = rad;
        }

        const std::uint32_t dst = warpAppend(outCount, isHit);
        if (!isHit) continue;

        const SurfaceShade sh = shadeSurfaceHit(fp, scene, hit, idx, v3(thr.x, thr.y, thr.z), sunDir);
        const V3           p = sh.p, wi = sh.wi, nextThroughput = sh.nextThroughput, contribution = sh.contribution;

        out.originPix[dst] = make_float4(p.x, p.y, p.z, oPix.w);
        out.direction[dst] = make_float4(wi.x, wi.y, wi.z, 0.0f);
        out.throughput[dst] = make_float4(nextThroughput.x, nextThroughput.y, nextThroughput.z, 0.0f);
        out.contribution[dst] = make_float4(contribution.x, contribution.y, contribution.z, 0.0f);
    }
}

// ---------------------------------------------------------------------------------------------
// The traversal launch.  Work items [0, numClosest) are closest-hit rays of `closestQueue`
// (rayIntersectBvh(ray, T_MAX, &hit), wgsl:190), items [numClosest, numClosest + numShadow) are the shadow
// rays of `shadowQueue` (shadowRay(Ray(p, lightDirection), T_MAX), wgsl:202, whose result is folded into the
// NEE term of rayColor:203).  After the k_shade of bounce b both the shadow rays of bounce b and the
// closest-hit rays of bounce b + 1 are known (same queue, different directions), so they share one launch:
// 9 traversal launches per 8-bounce frame instead of 16, and one kernel tail instead of two.
struct TraceIO : CursorSource
{
    static constexpr bool HANDS_OVER_STRAGGLERS = true; // at run time: stragglers.evictMax != 0
    // k_trace goes on with the whole warp once the queue is dry and the warp holds a single ray (run time: tuning.walkInPlace)
    static constexpr bool WALKS_LAST_RAY_WITH_WARP = true;
    __device__ __forceinline__ bool tailPhase(const bool queueDry) const { return queueDry && scene.tuning.walkInPlace != 0u; }
    const StragglerBuffer stragglers;
    const FrameParams&  fp;
    const SceneDevice&  scene;
    const PathQueue     closestQueue;
    const std::uint32_t numClosest;
    HitRecord*          hits;
    const PathQueue     shadowQueue;
    float4*             radiance;
    const V3            sunDir;
    std::uint32_t*      blockStats; // shared: closest {rays, nodes, tris}, shadow {rays, nodes, tris}

    __device__ __forceinline__ bool fetch(std::uint32_t& id, V3& o, V3& d, float& tmax, bool& anyHit) const
    {
        tmax = 10000.0f; // T_MAX, wgsl:73
        anyHit = id >= numClosest;
        if (anyHit) id -= numClosest; // the queue entry this work item traces
        if (!anyHit)
        {
            const float4 oo = closestQueue.originPix[id];
            const float4 dd = closestQueue.direction[id];
            o = v3(oo.x, oo.y, oo.z), d = v3(dd.x, dd.y, dd.z);
            return true;
        }
        const float4 oPix = shadowQueue.originPix[id];
        o = v3(oPix.x, oPix.y, oPix.z);
        d = sunSampleDirection(fp, scene, __float_as_uint(oPix.w), sunDir);
        return true;
    }
    __device__ __forceinline__ bool finish(
        const std::uint32_t i, const bool didHit, const HitRecord& hit, const std::uint32_t visited, const std::uint32_t tested, const bool anyHit,
        V3&, V3&, float&, bool&) const
    {
        std::uint32_t* st = blockStats + (anyHit ? 3 : 0);
        atomicAdd(st + 0, 1u);
        atomicAdd(st + 1, visited);
        atomicAdd(st + 2, tested);
        if (!anyHit)
        {
            hits[i] = hit;
            return false;
        }
        const std::uint32_t j = i;
        const std::uint32_t idx = __float_as_uint(shadowQueue.originPix[j].w);
        const float         vis = didHit ? 0.0f : 1.0f; // "Returns 1.0 if no forward intersections, 0.0 otherwise"
        const float4        c = shadowQueue.contribution[j];
        float4              rad = radiance[idx];
        if (fp.deferred)
        {
            // radiance += throughput * lightSample(...) (deferred_renderer_lighting_pass.wgsl:153,182); c.w != 0 marks the
            // primary surface, whose throughput is exactly 1
            const float4 t = c.w != 0.0f ? make_float4(1.f, 1.f, 1.f, 0.f) : shadowQueue.throughput[j];
            rad.x += t.x * (c.x * vis * fp.solarInvPdf);
            rad.y += t.y * (c.y * vis * fp.solarInvPdf);
            rad.z += t.z * (c.z * vis * fp.solarInvPdf);
            radiance[idx] = rad;
            return false;
        }
        rad.x += c.x * vis * fp.solarInvPdf;
        rad.y += c.y * vis * fp.solarInvPdf;
        rad.z += c.z * vis * fp.solarInvPdf;
        radiance[idx] = rad;
        return false;
    }
};

#ifndef RF_TRACE_MIN_BLOCKS
#define RF_TRACE_MIN_BLOCKS 4
#endif
template<int VARIANT, int BLOCK, int STACK = RF_STACK_SIZE>
__global__ void __launch_bounds__(BLOCK, RF_TRACE_MIN_BLOCKS * 256 / BLOCK) k_trace(
    const FrameParams    fp,
    const SceneDevice    scene,
    const PathQueue      closestQueue,
    const std::uint32_t* __restrict__ closestCount, // nullptr: no closest-hit rays in this launch
    HitRecord*           hits,
    const PathQueue      shadowQueue,
    const std::uint32_t* __restrict__ shadowCount,  // nullptr: no shadow rays in this launch
    float4*              radiance,
    std::uint32_t*       fetchCursor,
    const StragglerBuffer stragglers,               // evictMax == 0: every ray ends in this launch
    unsigned long long*  stats)
{
    __shared__ std::uint32_t blockStats[6];
    if (threadIdx.x < 6) blockStats[threadIdx.x] = 0u;
    __syncthreads();
    const std::uint32_t numClosest = closestCount ? *closestCount : 0u;
    const std::uint32_t numShadow = shadowCount ? *shadowCount : 0u;
    TraceIO io{{fetchCursor, numClosest + numShadow}, stragglers, fp, scene, closestQueue, numClosest, hits, shadowQueue, radiance,
                  v3(fp.sky.sun_direction), blockStats};
    // (the stacks are declared here so that a warp's own stack memory can serve as the scratch of its last ray's walk)
    __shared__ __align__(128) std::uint32_t stackMemory[STACK * BLOCK];
    static_assert(STACK * 128 >= static_cast<int>(sizeof(StragglerWindowShared)) && offsetof(StragglerWindowShared, stack) == 8 * 128, "the walk's scratch is the warp's own stack memory");
    WarpRay leftover;
    traceRays<2, VARIANT, BLOCK, TraceIO, STACK, true>(scene.nodes, scene.tris, scene.ordered, scene.tuning, io, stackMemory, &leftover);
    if (leftover.state != 0)
    {
        // the queue is dry and this warp held one ray: all 32 lanes walk it through 32-node windows (straggler.cuh), in place
        std::uint32_t parity = 0;
        traceWarpRay<STRAGGLER_DIRECT>(scene.nodes, scene.tris, leftover, *reinterpret_cast<StragglerWindowShared*>(stackMemory + (threadIdx.x >> 5) * (STACK * 32)), parity, io);
    }
    __syncthreads();
    if (threadIdx.x < 6 && blockStats[threadIdx.x] != 0u)
    {
        const int slot[6] = {STAT_CLOSEST_RAYS, STAT_CLOSEST_NODES, STAT_CLOSEST_TRIS, STAT_SHADOW_RAYS, STAT_SHADOW_NODES, STAT_SHADOW_TRIS};
        atomicAdd(&stats[slot[threadIdx.x]], static_cast<unsigned long long>(blockStats[threadIdx.x]));
        if (threadIdx.x == 1 || threadIdx.x == 4) atomicAdd(&stats[STAT_RECORDS], static_cast<unsigned long long>(blockStats[threadIdx.x])); // one node per visit
    }
}

// The traversal launch over child-pair records (traversal_pairs.cuh): same work items, same IO, same results.
template<int VARIANT, int BLOCK>
__global__ void __launch_bounds__(BLOCK, RF_TRACE_MIN_BLOCKS * 256 / BLOCK) k_trace_pairs(
    const __grid_constant__ FrameParams fp,
    const __grid_constant__ SceneDevice scene,
    const PathQueue      closestQueue,
    const std::uint32_t* __restrict__ closestCount, // nullptr: no closest-hit rays in this launch
    HitRecord*           hits,
    const PathQueue      shadowQueue,
    const std::uint32_t* __restrict__ shadowCount,  // nullptr: no shadow rays in this launch
    float4*              radiance,
    std::uint32_t*       fetchCursor,
    unsigned long long*  stats)
{
    __shared__ std::uint32_t blockStats[7];
    if (threadIdx.x < 7) blockStats[threadIdx.x] = 0u;
    __syncthreads();
    const std::uint32_t numClosest = closestCount ? *closestCount : 0u;
    const std::uint32_t numShadow = shadowCount ? *shadowCount : 0u;
    TraceIO io{{fetchCursor, numClosest + numShadow}, StragglerBuffer{nullptr, nullptr, 0u, 0u, 0u}, fp, scene, closestQueue, numClosest, hits, shadowQueue,
                  radiance, v3(fp.sky.sun_direction), blockStats};
    std::uint32_t records = 0;
    traceRaysPairs<2, VARIANT, BLOCK>(scene.pairs, scene.tris, scene.ordered, scene.tuning, io, records);
    records = __reduce_add_sync(0xFFFFFFFFu, records);
    if (laneId() == 0u) atomicAdd(&blockStats[6], records);
    __syncthreads();
    if (threadIdx.x < 7 && blockStats[threadIdx.x] != 0u)
    {
        const int slot[7] = {STAT_CLOSEST_RAYS, STAT_CLOSEST_NODES, STAT_CLOSEST_TRIS, STAT_SHADOW_RAYS, STAT_SHADOW_NODES, STAT_SHADOW_TRIS, STAT_RECORDS};
        atomicAdd(&stats[slot[threadIdx.x]], static_cast<unsigned long long>(blockStats[threadIdx.x]));
    }
}

// The follow-up of a k_trace launch that handed over its stragglers: one WARP per ray (straggler.cuh) — the warp
// loads 32 consecutive nodes per memory round trip, tests them lane-parallel and walks the ray through the
// results, ~3x faster per ray than a lone lane of the persistent loop.  Warps take records one at a time.
constexpr int STRAGGLER_BLOCK_THREADS = STRAGGLER_WARPS_PER_BLOCK * 32;
template<int WINDOW_MODE>
__global__ void __launch_bounds__(STRAGGLER_BLOCK_THREADS) k_trace_stragglers(
    const FrameParams     fp,
    const SceneDevice     scene,
    const PathQueue       closestQueue,
    HitRecord*            hits,
    const PathQueue       shadowQueue,
    float4*               radiance,
    std::uint32_t*        fetchCursor,
    const StragglerBuffer stragglers,
    unsigned long long*   stats)
{
    __shared__ std::uint32_t       blockStats[6];
    __shared__ StragglerWarpShared warpShared[STRAGGLER_WARPS_PER_BLOCK];
    const std::uint32_t numRecords = min(*stragglers.count, stragglers.capacity);
    if (blockIdx.x * STRAGGLER_WARPS_PER_BLOCK >= numRecords) return; // more warps than rays: nothing for this block
    if (threadIdx.x < 6) blockStats[threadIdx.x] = 0u;
    __syncthreads();
    TraceIO io{{fetchCursor, numRecords}, stragglers, fp, scene, closestQueue, 0u, hits, shadowQueue, radiance, v3(fp.sky.sun_direction), blockStats};
    StragglerWarpShared& mySh = warpShared[threadIdx.x >> 5];
    std::uint32_t        barrierParity = 0;
    if (WINDOW_MODE == STRAGGLER_BULK)
    {
        if (laneId() == 0u)
        {
            mbarrierInit(sharedAddress(&mySh.barrier[0]), 1u);
            mbarrierInit(sharedAddress(&mySh.barrier[1]), 1u);
            fenceProxyAsync(); // the barriers must be visible to the async proxy before the first copy names them
        }
        __syncwarp();
    }
#ifdef RF_TRACE_TIMELINE
    const unsigned long long tlStart = globalTimerNs();
    std::uint32_t            tlRays = 0;
#endif
    while (true)
    {
        std::uint32_t idx = 0;
        if (laneId() == 0u) idx = atomicAdd(fetchCursor, 1u);
        idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
        if (idx >= numRecords) break;
        traceStragglerWarp<WINDOW_MODE>(scene.nodes, scene.tris, stragglers.records + idx, mySh, barrierParity, io);
#ifdef RF_TRACE_TIMELINE
        ++tlRays;
#endif
    }
#ifdef RF_TRACE_TIMELINE
    if (laneId() == 0u && g_timeline != nullptr)
    {
        const std::uint32_t at = atomicAdd(&g_timelineCount, 1u);
        if (at < g_timelineCap) g_timeline[at] = TimelineRecord{io.timelineTag(), tlStart, tlStart, globalTimerNs(), tlRays, 0u, 0u, 0u};
    }
#endif
    __syncthreads();
    if (threadIdx.x < 6 && blockStats[threadIdx.x] != 0u)
    {
        const int slot[6] = {STAT_CLOSEST_RAYS, STAT_CLOSEST_NODES, STAT_CLOSEST_TRIS, STAT_SHADOW_RAYS, STAT_SHADOW_NODES, STAT_SHADOW_TRIS};
        atomicAdd(&stats[slot[threadIdx.x]], static_cast<unsigned long long>(blockStats[threadIdx.x]));
        if (threadIdx.x == 1 || threadIdx.x == 4) atomicAdd(&stats[STAT_RECORDS], static_cast<unsigned long long>(blockStats[threadIdx.x]));
    }
}

// ---------------------------------------------------------------------------------------------
// `peerImage` (multi-GPU, may be null): the HDR buffer of the root rank, mapped over NVLink peer memory.  Every pixel has
// exactly one owner, so writing the owner's accumulated value there IS the per-frame exchange: accumulate and exchange in
// one kernel, 16 B per owned pixel over NVLink, instead of a sum-reduce of W x H x 16 B full of zeros.
__global__ void __launch_bounds__(BLOCK_THREADS) k_accumulate(
    const FrameParams fp,
    const std::uint32_t* __restrict__ ownedTiles,
    const float4* __restrict__ radiance,
    float4*           image,
    float4*           peerImage,
    const bool        restart) // first sample of an accumulation: imageBuffer[idx] = vec3(0f) (fsMain:45-47)
{
    const std::uint32_t total = fp.numOwnedTiles * TILE_PIXELS;
    for (std::uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += gridDim.x * blockDim.x)
    {
        std::uint32_t px, py;
        if (!slotToPixel(fp, ownedTiles, slot, px, py)) continue;
        const std::uint32_t idx = py * fp.width + px;
        const float4        r = radiance[idx];
        float4              im = restart ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : image[idx];
        im.x += r.x, im.y += r.y, im.z += r.z;
        image[idx] = im;
        if (peerImage) peerImage[idx] = im;
    }
}

// acesFilmic, wgsl:278-285.
__device__ __forceinline__ float acesFilmic(const float x)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    const float y = __fdiv_rn(x * (a * x + b), x * (c * x + d) + e);
    return fminf(fmaxf(y, 0.0f), 1.0f);
}

__global__ void __launch_bounds__(BLOCK_THREADS) k_display(
    const std::uint32_t numPixels,
    const float4* __restrict__ image,
    const float       accumulated,
    const float       exposure,
    std::uint32_t*    outBgra)
{
    for (std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += gridDim.x * blockDim.x)
    {
        const float4 im = image[i];
        const float  rgb[3] = {
            powf(acesFilmic(exposure * __fdiv_rn(im.x, accumulated)), 1.0f / 2.2f),
            powf(acesFilmic(exposure * __fdiv_rn(im.y, accumulated)), 1.0f / 2.2f),
            powf(acesFilmic(exposure * __fdiv_rn(im.z, accumulated)), 1.0f / 2.2f)};
        std::uint32_t q[3];
        for (int c = 0; c < 3; ++c)
        {
            float v = rgb[c];
            v = (v != v) ? 0.0f : fminf(fmaxf(v, 0.0f), 1.0f); // unorm conversion: NaN -> 0, clamp, round
            q[c] = static_cast<std::uint32_t>(v * 255.0f + 0.5f);
        }
        outBgra[i] = q[2] | (q[1] << 8) | (q[0] << 16) | (255u << 24);
    }
}

// ---------------------------------------------------------------------------------------------
// Scene upload: repack the reference layouts into the traversal layouts (see traversal.cuh).
__global__ void k_pack_nodes(const rf_bvh_node* __restrict__ src, const std::uint64_t n, PackedNode* dst)
{
    for (std::uint64_t i = blockIdx.x * static_cast<std::uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<std::uint64_t>(gridDim.x) * blockDim.x)
    {
        const rf_bvh_node   s = src[i];
        const bool          leaf = s.triangle_count > 0u;
        const std::uint32_t A = leaf ? s.triangles_offset : s.second_child_offset;
        const std::uint32_t B = leaf ? ((s.triangle_count << 2) | 3u) : s.split_axis;
        dst[i] = PackedNode{s.aabb_min[0], s.aabb_min[1], s.aabb_min[2], s.aabb_max[0], s.aabb_max[1], s.aabb_max[2], A, B};
    }
}

// `src` = triangles with `strideFloats` floats per vertex (4 for PositionAttribute, 3 for Positions).
__global__ void k_pack_triangles(const float* __restrict__ src, const int strideFloats, const std::uint64_t n, float4* dst)
{
    for (std::uint64_t i = blockIdx.x * static_cast<std::uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<std::uint64_t>(gridDim.x) * blockDim.x)
    {
        const float* t = src + i * 3 * strideFloats;
        const V3     v0 = v3(t), v1 = v3(t + strideFloats), v2 = v3(t + 2 * strideFloats);
        const V3     e1 = v1 - v0, e2 = v2 - v0;
        const V3     nrm = normalize(cross(e1, e2)); // wgsl:513 / ray_intersection.cpp:80
        dst[TRI_STRIDE * i + 0] = make_float4(v0.x, v0.y, v0.z, e1.x);
        dst[TRI_STRIDE * i + 1] = make_float4(e1.y, e1.z, e2.x, e2.y);
        dst[TRI_STRIDE * i + 2] = make_float4(e2.z, nrm.x, nrm.y, nrm.z);
        dst[TRI_STRIDE * i + 3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// ---------------------------------------------------------------------------------------------
// bvh-visualizer pixel loop (bvh-visualizer/main.cpp:60-78) and the batched rayIntersectBvh.
struct VisualizerIO : CursorSource
{
    const rf_camera     camera;
    const std::uint32_t width, height, blocksX;
    const float         rayTMax;
    std::uint32_t*      outNodes;
    // ray i -> pixel: 8x4 pixel blocks (one per 32 consecutive rays), blocks in row-major order
    __device__ __forceinline__ bool pixel(const std::uint32_t i, std::uint32_t& px, std::uint32_t& py) const
    {
        const std::uint32_t block = i >> 5, lane = i & 31u;
        px = (block % blocksX) * 8u + (lane & 7u);
        py = (block / blocksX) * 4u + (lane >> 3);
        return px < width && py < height;
    }
    __device__ __forceinline__ bool fetch(std::uint32_t& i, V3& o, V3& d, float& tmax, bool&) const
    {
        std::uint32_t j, row;
        if (!pixel(i, j, row)) return false;
        const float u = __fdiv_rn(static_cast<float>(j), static_cast<float>(width));
        const float v = 1.0f - __fdiv_rn(static_cast<float>(row + 1u), static_cast<float>(height));
        // generateCameraRay, camera.cpp:44-51
        o = v3(camera.origin);
        d = normalize(((v3(camera.lower_left_corner) + v3(camera.horizontal) * u) + v3(camera.vertical) * v) - o);
        tmax = rayTMax;
        return true;
    }
    __device__ __forceinline__ bool finish(const std::uint32_t i, bool, const HitRecord&, const std::uint32_t visited, std::uint32_t, bool, V3&, V3&, float&, bool&) const
    {
        std::uint32_t j, row;
        pixel(i, j, row);
        outNodes[row * width + j] = visited;
        return false;
    }
};

__global__ void __launch_bounds__(TRACE_BLOCK_THREADS) k_visualizer(
    const PackedNode* __restrict__ nodes,
    const float4* __restrict__ tris,
    const bool          ordered,
    const TraceTuning   tuning,
    const rf_camera     camera,
    const std::uint32_t width,
    const std::uint32_t height,
    const float         rayTMax,
    std::uint32_t*      cursor,
    std::uint32_t*      outNodes)
{
    const std::uint32_t blocksX = (width + 7u) / 8u, blocksY = (height + 3u) / 4u;
    VisualizerIO        io{{cursor, blocksX * blocksY * 32u}, camera, width, height, blocksX, rayTMax, outNodes};
    traceRays<0, TRACE_DEFAULT_VARIANT, TRACE_BLOCK_THREADS>(nodes, tris, ordered, tuning, io);
}

__global__ void __launch_bounds__(TRACE_BLOCK_THREADS) k_visualizer_pairs(
    const __grid_constant__ PairSceneDevice pairs,
    const float4* __restrict__ tris,
    const bool          ordered,
    const TraceTuning   tuning,
    const rf_camera     camera,
    const std::uint32_t width,
    const std::uint32_t height,
    const float         rayTMax,
    std::uint32_t*      cursor,
    std::uint32_t*      outNodes)
{
    const std::uint32_t blocksX = (width + 7u) / 8u, blocksY = (height + 3u) / 4u;
    VisualizerIO        io{{cursor, blocksX * blocksY * 32u}, camera, width, height, blocksX, rayTMax, outNodes};
    std::uint32_t       records = 0;
    traceRaysPairs<0, PAIR_DEFAULT_VARIANT, TRACE_BLOCK_THREADS>(pairs, tris, ordered, tuning, io, records);
}

struct BatchIO : CursorSource
{
    const float*   rays;
    const float4*  tris;
    const float    rayTMax;
    std::uint8_t*  outHit;
    float4*        outPT;
    std::uint32_t* outNodes;
    __device__ __forceinline__ bool fetch(std::uint32_t& i, V3& o, V3& d, float& tmax, bool&) const
    {
        const float* r = rays + 6ull * i;
        o = v3(r[0], r[1], r[2]), d = v3(r[3], r[4], r[5]);
        tmax = rayTMax;
        return true;
    }
    __device__ __forceinline__ bool finish(const std::uint32_t i, const bool didHit, const HitRecord& hit, const std::uint32_t visited, std::uint32_t, bool, V3&, V3&, float&, bool&) const
    {
        if (outHit) outHit[i] = didHit ? 1 : 0;
        if (outPT)
        {
            float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
            if (didHit)
            {
                const V3 p = hitPoint(tris, hit);
                pt = make_float4(p.x, p.y, p.z, hit.t);
            }
            outPT[i] = pt;
        }
        if (outNodes) outNodes[i] = visited;
        return false;
    }
};

__global__ void __launch_bounds__(TRACE_BLOCK_THREADS) k_intersect_batch(
    const PackedNode* __restrict__ nodes,
    const float4* __restrict__ tris,
    const bool          ordered,
    const TraceTuning   tuning,
    const float* __restrict__ rays,
    const std::uint32_t numRays,
    const float         rayTMax,
    std::uint32_t*      cursor,
    std::uint8_t*       outHit,
    float4*             outPT,
    std::uint32_t*      outNodes)
{
    BatchIO io{{cursor, numRays}, rays, tris, rayTMax, outHit, outPT, outNodes};
    traceRays<0, TRACE_DEFAULT_VARIANT, TRACE_BLOCK_THREADS>(nodes, tris, ordered, tuning, io);
}

__global__ void __launch_bounds__(TRACE_BLOCK_THREADS) k_intersect_batch_pairs(
    const __grid_constant__ PairSceneDevice pairs,
    const float4* __restrict__ tris,
    const bool          ordered,
    const TraceTuning   tuning,
    const float* __restrict__ rays,
    const std::uint32_t numRays,
    const float         rayTMax,
    std::uint32_t*      cursor,
    std::uint8_t*       outHit,
    float4*             outPT,
    std::uint32_t*      outNodes)
{
    BatchIO       io{{cursor, numRays}, rays, tris, rayTMax, outHit, outPT, outNodes};
    std::uint32_t records = 0;
    traceRaysPairs<0, PAIR_DEFAULT_VARIANT, TRACE_BLOCK_THREADS>(pairs, tris, ordered, tuning, io, records);
}
} // namespace rfb200

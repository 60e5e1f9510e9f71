// BVH traversal + Moeller-Trumbore for sm_100a — the device twin of
//   rayIntersectBvh   reference_path_tracer.wgsl:371-429  / common/ray_intersection.cpp:138-213
//   shadowRay         reference_path_tracer.wgsl:323-368
//   rayIntersectAabb  wgsl:448-475 / ray_intersection.cpp:101-136
//   rayIntersectTriangle wgsl:478-521 / ray_intersection.cpp:38-90
//
// The traversal ORDER and every fp32 operation are those of the reference (strict IEEE, no FMA:
// this file is compiled with -fmad=false), so nodesVisited is bit-exact and the closest hit — ties
// included — is the one the reference finds.  What differs is the memory layout the data is pulled
// through (DESIGN.md "Data layout in HBM"):
//
//   node  (32 B = 2 x float4, one 32-byte sector instead of the reference's 48 B / two sectors)
//         q0 = (min.x, min.y, min.z, max.x)   q1 = (max.y, max.z, A, B)
//         interior: A = secondChildOffset, B = splitAxis (0..2)
//         leaf:     A = trianglesOffset,   B = (triangleCount << 2) | 3
//   tri   (48 B = 3 x float4)  (v0.xyz, e1.x) (e1.yz, e2.xy) (e2.z, n.xyz)
//         e1 = v1 - v0, e2 = v2 - v0, n = normalize(cross(e1, e2)) precomputed once at upload with the
//         same fp32 operations the reference performs per test (so results are unchanged).
#pragma once

#include "rf_vec.h"

#include <cstdint>
#include <cuda_runtime.h>

namespace rfb200
{
constexpr float         RF_EPSILON = 0.00001f; // wgsl:66, ray_intersection.cpp:44
constexpr int           RF_STACK_SIZE = 32;    // wgsl:327,375; ray_intersection.cpp:148
constexpr std::uint32_t RF_NO_HIT = 0xFFFFFFFFu;

struct HitRecord
{
    std::uint32_t tri; // RF_NO_HIT on miss
    float         u, v, t;
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

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
    const float4 a = ldg4(tris + 3 * tri + 0);
    const float4 b = ldg4(tris + 3 * tri + 1);
    const float4 c = ldg4(tris + 3 * tri + 2);
    const V3     v0 = v3(a.x, a.y, a.z);
    const V3     e1 = v3(a.w, b.x, b.y);
    const V3     e2 = v3(b.z, b.w, c.x);

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

// Iterative pre-order traversal with an explicit 32-entry stack.  ANY_HIT = shadowRay semantics
// (constant rayTMax, return on the first accepted triangle); otherwise closest hit with shrinking
// tmax.  `nodesVisited` counts loop iterations exactly like ray_intersection.cpp:158.
template<bool ANY_HIT>
__device__ __forceinline__ bool traverseBvh(
    const float4* __restrict__ nodes,
    const float4* __restrict__ tris,
    const V3       o,
    const V3       d,
    float          tmax,
    HitRecord&     hit,
    std::uint32_t& nodesVisited,
    std::uint32_t& trianglesTested)
{
    // rayAabbIntersector, wgsl:438-445 / ray_intersection.cpp:92-99
    const float ix = __fdiv_rn(1.0f, d.x), iy = __fdiv_rn(1.0f, d.y), iz = __fdiv_rn(1.0f, d.z);
    const bool  negX = ix < 0.0f, negY = iy < 0.0f, negZ = iz < 0.0f;

    std::uint32_t stack[RF_STACK_SIZE];
    int           sp = 0;
    std::uint32_t cur = 0;
    bool          didHit = false;
    hit.tri = RF_NO_HIT;

    while (true)
    {
        ++nodesVisited;
        const float4 q0 = ldg4(nodes + 2 * cur);
        const float4 q1 = ldg4(nodes + 2 * cur + 1);

        // rayIntersectAabb: bounds[dirNeg] / bounds[1 - dirNeg] selection, (b - o) * invDir, the two
        // early-outs and std::max/std::min operand order (NaN-propagating) are the reference's.
        const float loX = negX ? q0.w : q0.x, hiX = negX ? q0.x : q0.w;
        const float loY = negY ? q1.x : q0.y, hiY = negY ? q0.y : q1.x;
        const float loZ = negZ ? q1.y : q0.z, hiZ = negZ ? q0.z : q1.y;
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
        boxHit = boxHit && (tmin < tmax) && (tmx > 0.0f);

        const std::uint32_t A = __float_as_uint(q1.z);
        const std::uint32_t B = __float_as_uint(q1.w);

        if (boxHit && B < 3u)
        {
            // interior: visit the near child first by the sign of invDir[splitAxis]
            const bool neg = (B == 0u) ? negX : ((B == 1u) ? negY : negZ);
            stack[sp++] = neg ? cur + 1u : A;
            cur = neg ? A : cur + 1u;
            continue;
        }
        if (boxHit)
        {
            const std::uint32_t count = B >> 2;
            for (std::uint32_t k = 0; k < count; ++k)
            {
                ++trianglesTested;
                float u, v, t;
                if (intersectTriangle(tris, A + k, o, d, tmax, u, v, t))
                {
                    if (ANY_HIT) return true;
                    tmax = t;
                    didHit = true;
                    hit.tri = A + k, hit.u = u, hit.v = v, hit.t = t;
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return didHit;
}

// offsetRay, wgsl:523-544 / ray_intersection.cpp:17-35.
__device__ __forceinline__ float offsetRayComponent(const float p, const float n)
{
    const float ORIGIN = 1.0f / 32.0f;
    const float FLOAT_SCALE = 1.0f / 65536.0f;
    const float INT_SCALE = 256.0f;
    const int   off = __float2int_rz(INT_SCALE * n);
    const float po = __int_as_float(__float_as_int(p) + ((p < 0.0f) ? -off : off));
    return (fabsf(p) < ORIGIN) ? (p + FLOAT_SCALE * n) : po;
}

// Hit point of an accepted triangle: p = v0 + u*e1 + v*e2, then offsetRay(p, n) (wgsl:509-516).
__device__ __forceinline__ V3 hitPoint(const float4* __restrict__ tris, const HitRecord& hit)
{
    const float4 a = ldg4(tris + 3 * hit.tri + 0);
    const float4 b = ldg4(tris + 3 * hit.tri + 1);
    const float4 c = ldg4(tris + 3 * hit.tri + 2);
    const V3     v0 = v3(a.x, a.y, a.z);
    const V3     e1 = v3(a.w, b.x, b.y);
    const V3     e2 = v3(b.z, b.w, c.x);
    const V3     n = v3(c.y, c.z, c.w);
    const V3     p = (v0 + hit.u * e1) + hit.v * e2;
    return v3(offsetRayComponent(p.x, n.x), offsetRayComponent(p.y, n.y), offsetRayComponent(p.z, n.z));
}
} // namespace rfb200

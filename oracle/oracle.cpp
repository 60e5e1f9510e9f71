// TEST INFRASTRUCTURE — the parity oracle.  NOT part of the product path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// oracle/liboracle.so; rayfinder_b200/ never does.
//
// A scalar CPU restatement, function by function, of the reference's render path:
//   Oracle A  common/ray_intersection.cpp:17-213 (CPU rayIntersectBvh + BvhStats) and the pixel loop
//             of bvh-visualizer/main.cpp:60-78, common/camera.cpp:7-51
//   Oracle B  pt/reference_path_tracer.wgsl:34-616 (fsMain minus the display transform, plus the
//             display transform separately) with the host-side packing of
//             pt/reference_path_tracer.cpp:168-184 (blue noise /255) and :209-270 (texture descriptors)
//   Oracle C  pt/deferred_renderer_lighting_pass.wgsl:96-236,497-519 (lighting pass: main, worldFromUv, surfaceColor,
//             lightSample, skyRadiance with the solar disk, offsetPosition) and
//             pt/deferred_renderer_resolve_pass.wgsl:34-50 (moving average), on a caller-supplied G-buffer
// Strict fp32: compiled with -ffp-contract=off, every expression in the reference's operand order;
// vector helpers follow glm 0.9.9.8's scalar formulas (dot = (x+y)+z, normalize = v * (1/sqrt(dot))).
// WGSL builtins whose rounding the WGSL spec leaves open (dot/normalize/mat*vec summation order,
// cos/sin/acos/exp/pow accuracy) are pinned to those formulas and to this box's libm.
//
// Pinning status:
//   * Oracle A is checked against the reference's own compiled translation units
//     (oracle/_ref/libref_oracle.so: bit-exact t, p, nodesVisited; tests/test_oracle_vs_ref.py) and
//     against the reference's known-answer/property tests (tests/aabb.cpp:61-132,
//     tests/intersection.cpp:9-28, tests/bvh.cpp:34-102).
//   * Oracle B: the WGSL cannot be executed here (needs Dawn + a GPU API) and the reference has no
//     test or golden image for it => RADIANCE PARITY IS UNPINNED against a real
//     reference build.  Its traversal/triangle arithmetic is the pinned Oracle A code; its sky
//     evaluation is checked against the reference's sky_state_radiance (hw_skymodel.c:182-223).
//   * Oracle C: same situation as Oracle B (WGSL compute pass, no reference test or golden image) => PARITY UNPINNED
//     against a real reference build; it shares Oracle A's pinned traversal/triangle code and Oracle B's helpers.
#include <atomic>
#include <bit>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
struct vec3
{
    float x, y, z;
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3  cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline vec3  normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline vec3  load3(const float* p) { return {p[0], p[1], p[2]}; }

struct Ray
{
    vec3 origin, direction;
};

// ---- layouts (byte-identical to the reference structs) -----------------------------------------
struct BvhNode // common/bvh.hpp:13-21
{
    float         min[3], pad0, max[3], pad1;
    std::uint32_t trianglesOffset, secondChildOffset, triangleCount, splitAxis;
};
static_assert(sizeof(BvhNode) == 48);
struct VertexAttributes // pt-format/vertex_attributes.hpp:17-35
{
    float         n0[3], pad0, n1[3], pad1, n2[3], pad2;
    float         uv0[2], uv1[2], uv2[2];
    std::uint32_t textureIdx, pad3;
};
static_assert(sizeof(VertexAttributes) == 80);
struct Camera // common/camera.hpp:10-21
{
    float origin[3], lowerLeftCorner[3], horizontal[3], vertical[3], up[3], right[3], lensRadius;
};
static_assert(sizeof(Camera) == 76);
struct SkyState // pt/aligned_sky_state.hpp:34-41
{
    float params[27], skyRadiances[3], solarRadiances[3], padding1[3], sunDirection[3], padding2;
};
static_assert(sizeof(SkyState) == 160);

constexpr float EPSILON = 0.00001f; // wgsl:66
constexpr float PI = 3.1415927f;    // wgsl:68
constexpr float FRAC_1_PI = 0.31830987f;
constexpr float T_MAX = 10000.0f; // wgsl:73

// ---- offsetRay: ray_intersection.cpp:17-35 / wgsl:523-544 ---------------------------------------
// The deferred renderer's offsetPosition (deferred_renderer_lighting_pass.wgsl:497-519) is the same function with
// FLOAT_SCALE 1/16384 and INT_SCALE 1024; `deferred` selects it.
vec3 offsetRay(vec3 p, vec3 n, bool deferred = false)
{
    const float ORIGIN = 1.0f / 32.0f;
    const float FLOAT_SCALE = deferred ? 1.0f / 16384.0f : 1.0f / 65536.0f;
    const float INT_SCALE = deferred ? 1024.0f : 256.0f;
    const int   ox = int(INT_SCALE * n.x), oy = int(INT_SCALE * n.y), oz = int(INT_SCALE * n.z);
    const vec3  po = {
        std::bit_cast<float>(std::bit_cast<int>(p.x) + (p.x < 0 ? -ox : ox)),
        std::bit_cast<float>(std::bit_cast<int>(p.y) + (p.y < 0 ? -oy : oy)),
        std::bit_cast<float>(std::bit_cast<int>(p.z) + (p.z < 0 ? -oz : oz))};
    return {
        std::fabs(p.x) < ORIGIN ? p.x + FLOAT_SCALE * n.x : po.x,
        std::fabs(p.y) < ORIGIN ? p.y + FLOAT_SCALE * n.y : po.y,
        std::fabs(p.z) < ORIGIN ? p.z + FLOAT_SCALE * n.z : po.z};
}

// ---- rayIntersectTriangle: ray_intersection.cpp:38-90 / wgsl:478-521 ----------------------------
struct TriangleHit
{
    vec3  p;
    vec3  b;
    float t;
};
bool rayIntersectTriangle(const Ray& ray, vec3 p0, vec3 p1, vec3 p2, float tmax, TriangleHit& hit, bool deferred = false)
{
    const vec3  e1 = p1 - p0;
    const vec3  e2 = p2 - p0;
    const vec3  h = cross(ray.direction, e2);
    const float det = dot(e1, h);
    if (det > -EPSILON && det < EPSILON) return false;
    const float invDet = 1.0f / det;
    const vec3  s = ray.origin - p0;
    const float u = invDet * dot(s, h);
    if (u < 0.0f || u > 1.0f) return false;
    const vec3  q = cross(s, e1);
    const float v = invDet * dot(ray.direction, q);
    if (v < 0.0f || u + v > 1.0f) return false;
    const float t = invDet * dot(e2, q);
    if (t > EPSILON && t < tmax)
    {
        const vec3 p = p0 + u * e1 + v * e2;
        const vec3 n = normalize(cross(e1, e2));
        hit.p = offsetRay(p, n, deferred);
        hit.b = {1.0f - u - v, u, v};
        hit.t = t;
        return true;
    }
    return false;
}

// ---- RayAabbIntersector / rayIntersectAabb: ray_intersection.cpp:92-136 / wgsl:431-475 ----------
struct RayAabbIntersector
{
    vec3          origin, invDir;
    std::uint32_t dirNeg[3];
    explicit RayAabbIntersector(const Ray& ray)
    {
        origin = ray.origin;
        invDir = {1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z};
        dirNeg[0] = invDir.x < 0.0f, dirNeg[1] = invDir.y < 0.0f, dirNeg[2] = invDir.z < 0.0f;
    }
};
inline float stdMax(float a, float b) { return (a < b) ? b : a; } // std::max / WGSL max operand order
inline float stdMin(float a, float b) { return (b < a) ? b : a; } // std::min / WGSL min operand order
bool rayIntersectAabb(const RayAabbIntersector& is, const float* bmin, const float* bmax, float rayTMax)
{
    const float* bounds[2] = {bmin, bmax};
    float        tmin = (bounds[is.dirNeg[0]][0] - is.origin.x) * is.invDir.x;
    float        tmax = (bounds[1 - is.dirNeg[0]][0] - is.origin.x) * is.invDir.x;
    const float  tymin = (bounds[is.dirNeg[1]][1] - is.origin.y) * is.invDir.y;
    const float  tymax = (bounds[1 - is.dirNeg[1]][1] - is.origin.y) * is.invDir.y;
    if ((tmin > tymax) || (tymin > tmax)) return false;
    tmin = stdMax(tymin, tmin);
    tmax = stdMin(tymax, tmax);
    const float tzmin = (bounds[is.dirNeg[2]][2] - is.origin.z) * is.invDir.z;
    const float tzmax = (bounds[1 - is.dirNeg[2]][2] - is.origin.z) * is.invDir.z;
    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    tmin = stdMax(tzmin, tmin);
    tmax = stdMin(tzmax, tmax);
    return (tmin < rayTMax) && (tmax > 0.0f);
}

struct Counters
{
    std::uint64_t closestRays = 0, shadowRays = 0, closestNodes = 0, closestTris = 0, shadowNodes = 0, shadowTris = 0;
    std::uint64_t paths = 0;
    // optional log of every ray rayColor traces: (origin, direction) + kind (0 closest-hit, 1 shadow); see oracle_frame_rays
    std::vector<float>*        rayLog = nullptr;
    std::vector<std::uint8_t>* rayKind = nullptr;
};

// Scene view used by both oracles.  `triStride` = floats per vertex (3: nlrs::Positions, 4:
// PositionAttribute).
struct SceneView
{
    const BvhNode*          nodes;
    const float*            tris;
    int                     triStride;
    const VertexAttributes* vattr;     // Oracle B only
    const std::uint32_t*    texDesc;   // (width, height, offset) triples
    std::uint32_t           numTextures;
    const std::uint32_t*    texels;
    std::uint64_t           numTexels;
    const float*            blueNoise; // vec2f per texel = u8 / 255 (reference_path_tracer.cpp:174-177)
    std::uint32_t           bnWidth, bnHeight;
    bool                    deferred = false; // the deferred renderer's lighting pass: its offsetPosition constants
};

struct Intersection // wgsl:158-163
{
    vec3          p, n;
    float         uv[2];
    std::uint32_t textureDescriptorIdx;
};

// ---- rayIntersectBvh: ray_intersection.cpp:138-213 (stats) and wgsl:371-429 (attributes) ---------
// One body serves both twins: the traversal is identical; `hit` (WGSL attributes) is filled only when
// `scene.vattr` is set, `outT`/`outP` always (the CPU Intersection{p, t}).
bool rayIntersectBvh(
    const SceneView& scene,
    const Ray&       ray,
    float            rayTMax,
    Intersection*    hit,
    float*           outT,
    vec3*            outP,
    std::uint64_t&   nodesVisited,
    std::uint64_t&   trisTested)
{
    const RayAabbIntersector intersector(ray);
    std::uint32_t            toVisitOffset = 0, currentNodeIdx = 0;
    std::uint32_t            nodesToVisit[32];
    bool                     didIntersect = false;
    float                    tmax = rayTMax;
    for (;;)
    {
        ++nodesVisited;
        const BvhNode& node = scene.nodes[currentNodeIdx];
        if (rayIntersectAabb(intersector, node.min, node.max, tmax))
        {
            if (node.triangleCount > 0u)
            {
                for (std::uint32_t idx = 0; idx < node.triangleCount; ++idx)
                {
                    const std::uint32_t triangleIdx = node.trianglesOffset + idx;
                    const float*        t = scene.tris + static_cast<std::size_t>(triangleIdx) * 3 * scene.triStride;
                    TriangleHit         trihit;
                    ++trisTested;
                    if (rayIntersectTriangle(ray, load3(t), load3(t + scene.triStride), load3(t + 2 * scene.triStride), tmax, trihit, scene.deferred))
                    {
                        tmax = trihit.t;
                        didIntersect = true;
                        if (outT) *outT = trihit.t;
                        if (outP) *outP = trihit.p;
                        if (hit && scene.vattr)
                        {
                            const vec3              b = trihit.b;
                            const VertexAttributes& vert = scene.vattr[triangleIdx];
                            hit->p = trihit.p;
                            hit->n = b.x * load3(vert.n0) + b.y * load3(vert.n1) + b.z * load3(vert.n2);
                            hit->uv[0] = b.x * vert.uv0[0] + b.y * vert.uv1[0] + b.z * vert.uv2[0];
                            hit->uv[1] = b.x * vert.uv0[1] + b.y * vert.uv1[1] + b.z * vert.uv2[1];
                            hit->textureDescriptorIdx = vert.textureIdx;
                        }
                    }
                }
                if (toVisitOffset == 0u) break;
                currentNodeIdx = nodesToVisit[--toVisitOffset];
            }
            else
            {
                if (intersector.dirNeg[node.splitAxis] == 1u)
                {
                    nodesToVisit[toVisitOffset++] = currentNodeIdx + 1u;
                    currentNodeIdx = node.secondChildOffset;
                }
                else
                {
                    nodesToVisit[toVisitOffset++] = node.secondChildOffset;
                    currentNodeIdx = currentNodeIdx + 1u;
                }
            }
        }
        else
        {
            if (toVisitOffset == 0u) break;
            currentNodeIdx = nodesToVisit[--toVisitOffset];
        }
    }
    return didIntersect;
}

// ---- shadowRay: wgsl:323-368 ------------------------------------------------------------------
float shadowRay(const SceneView& scene, const Ray& ray, float rayTMax, std::uint64_t& nodesVisited, std::uint64_t& trisTested)
{
    const RayAabbIntersector intersector(ray);
    std::uint32_t            toVisitOffset = 0, currentNodeIdx = 0;
    std::uint32_t            nodesToVisit[32];
    for (;;)
    {
        ++nodesVisited;
        const BvhNode& node = scene.nodes[currentNodeIdx];
        if (rayIntersectAabb(intersector, node.min, node.max, rayTMax))
        {
            if (node.triangleCount > 0u)
            {
                for (std::uint32_t idx = 0; idx < node.triangleCount; ++idx)
                {
                    const float* t = scene.tris + static_cast<std::size_t>(node.trianglesOffset + idx) * 3 * scene.triStride;
                    TriangleHit  trihit;
                    ++trisTested;
                    if (rayIntersectTriangle(ray, load3(t), load3(t + scene.triStride), load3(t + 2 * scene.triStride), rayTMax, trihit))
                    {
                        return 0.0f;
                    }
                }
                if (toVisitOffset == 0u) break;
                currentNodeIdx = nodesToVisit[--toVisitOffset];
            }
            else
            {
                if (intersector.dirNeg[node.splitAxis] == 1u)
                {
                    nodesToVisit[toVisitOffset++] = currentNodeIdx + 1u;
                    currentNodeIdx = node.secondChildOffset;
                }
                else
                {
                    nodesToVisit[toVisitOffset++] = node.secondChildOffset;
                    currentNodeIdx = currentNodeIdx + 1u;
                }
            }
        }
        else
        {
            if (toVisitOffset == 0u) break;
            currentNodeIdx = nodesToVisit[--toVisitOffset];
        }
    }
    return 1.0f;
}

// ---- WGSL helpers -------------------------------------------------------------------------------
inline float fract(float x) { return x - std::floor(x); } // WGSL fract

// pixarOnb, wgsl:310-319, returning the three columns.
struct Mat3
{
    vec3 c0, c1, c2;
};
inline vec3 mul(const Mat3& m, vec3 v) { return v.x * m.c0 + v.y * m.c1 + v.z * m.c2; }
Mat3        pixarOnb(vec3 n)
{
    const float s = (n.z >= 0.0f) ? 1.0f : -1.0f;
    const float a = -1.0f / (s + n.z);
    const float b = n.x * n.y * a;
    const vec3  u = {1.0f + s * n.x * n.x * a, s * b, -s * n.x};
    const vec3  v = {b, s + n.y * n.y * a, -n.y};
    return {u, v, n};
}

// directionInCone, wgsl:569-579
vec3 directionInCone(const float u[2], float cosThetaMax)
{
    const float cosTheta = 1.0f - u[0] * (1.0f - cosThetaMax);
    const float sinTheta = std::sqrt(1.0f - cosTheta * cosTheta);
    const float phi = 2.0f * PI * u[1];
    return {std::cos(phi) * sinTheta, std::sin(phi) * sinTheta, cosTheta};
}
// directionInCosineWeightedHemisphere, wgsl:583-592
vec3 directionInCosineWeightedHemisphere(const float u[2])
{
    const float phi = 2.0f * PI * u[1];
    const float sinTheta = std::sqrt(1.0f - u[0]);
    return {std::cos(phi) * sinTheta, std::sin(phi) * sinTheta, std::sqrt(u[0])};
}

// textureLookup, wgsl:553-565 (+ robust-buffer-access clamp of the texel index)
vec3 textureLookup(const SceneView& scene, std::uint32_t descIdx, const float uv[2])
{
    if (descIdx >= scene.numTextures) descIdx = scene.numTextures - 1u;
    const std::uint32_t width = scene.texDesc[3 * descIdx], height = scene.texDesc[3 * descIdx + 1], offset = scene.texDesc[3 * descIdx + 2];
    const float         u = fract(uv[0]);
    const float         v = fract(uv[1]);
    const std::uint32_t j = static_cast<std::uint32_t>(u * static_cast<float>(width));
    const std::uint32_t i = static_cast<std::uint32_t>(v * static_cast<float>(height));
    const std::uint32_t idx = i * width + j;
    std::uint64_t       at = static_cast<std::uint64_t>(offset) + idx;
    if (at >= scene.numTexels) at = scene.numTexels - 1;
    const std::uint32_t bgra = scene.texels[at];
    const vec3          srgb = {
        static_cast<float>((bgra >> 16u) & 0xffu) / 255.0f,
        static_cast<float>((bgra >> 8u) & 0xffu) / 255.0f,
        static_cast<float>(bgra & 0xffu) / 255.0f};
    return {std::pow(srgb.x, 2.2f), std::pow(srgb.y, 2.2f), std::pow(srgb.z, 2.2f)};
}

// skyRadiance, wgsl:248-275
float skyRadiance(const SkyState& sky, float theta, float gamma, std::uint32_t channel)
{
    const float  r = sky.skyRadiances[channel];
    const float* p = sky.params + 9u * channel;
    const float  cosGamma = std::cos(gamma);
    const float  cosGamma2 = cosGamma * cosGamma;
    const float  cosTheta = std::fabs(std::cos(theta));
    const float  expM = std::exp(p[4] * gamma);
    const float  rayM = cosGamma2;
    const float  mieMLhs = 1.0f + cosGamma2;
    const float  mieMRhs = std::pow(1.0f + p[8] * p[8] - 2.0f * p[8] * cosGamma, 1.5f);
    const float  mieM = mieMLhs / mieMRhs;
    const float  zenith = std::sqrt(cosTheta);
    const float  radianceLhs = 1.0f + p[0] * std::exp(p[1] / (cosTheta + 0.01f));
    const float  radianceRhs = p[2] + p[3] * expM + p[5] * rayM + p[6] * mieM + p[7] * zenith;
    const float  radianceDist = radianceLhs * radianceRhs;
    return r * radianceDist;
}

struct Constants
{
    float solarCosThetaMax, solarInvPdf;
};
Constants constants()
{
    // wgsl:78-83: f32 const-expressions.
    const float degreesToRadians = PI / 180.0f;
    const float terrestrialSolarRadius = 0.255f * degreesToRadians;
    const float c = std::cos(terrestrialSolarRadius);
    return {c, 2.0f * PI * (1.0f - c)};
}

// rayColor, wgsl:181-234
vec3 rayColor(
    const SceneView& scene,
    const SkyState&  sky,
    const float      blueNoise[2],
    Ray              ray,
    std::uint32_t    numBounces,
    Counters&        ctr)
{
    const Constants k = constants();
    vec3            radiance = {0.f, 0.f, 0.f};
    vec3            throughput = {1.f, 1.f, 1.f};
    const vec3      sunDirection = load3(sky.sunDirection);
    std::uint32_t   bounce = 1u;
    for (;;)
    {
        Intersection hit;
        ++ctr.closestRays;
        const auto logRay = [&ctr](const Ray& r, std::uint8_t kind) {
            if (!ctr.rayLog) return;
            const float six[6] = {r.origin.x, r.origin.y, r.origin.z, r.direction.x, r.direction.y, r.direction.z};
            ctr.rayLog->insert(ctr.rayLog->end(), six, six + 6);
            if (ctr.rayKind) ctr.rayKind->push_back(kind);
        };
        logRay(ray, 0);
        if (rayIntersectBvh(scene, ray, T_MAX, &hit, nullptr, nullptr, ctr.closestNodes, ctr.closestTris))
        {
            const vec3 albedo = textureLookup(scene, hit.textureDescriptorIdx, hit.uv);
            const vec3 p = hit.p;

            const vec3  lightDirection = mul(pixarOnb(sunDirection), directionInCone(blueNoise, k.solarCosThetaMax));
            const vec3  lightIntensity = {sky.solarRadiances[0], sky.solarRadiances[1], sky.solarRadiances[2]};
            const vec3  brdf = albedo * FRAC_1_PI;
            const vec3  reflectance = brdf * dot(hit.n, lightDirection);
            ++ctr.shadowRays;
            logRay(Ray{p, lightDirection}, 1);
            const float lightVisibility = shadowRay(scene, Ray{p, lightDirection}, T_MAX, ctr.shadowNodes, ctr.shadowTris);
            radiance = radiance + throughput * lightIntensity * reflectance * lightVisibility * k.solarInvPdf;

            if (bounce == numBounces) break;

            const vec3 wi = mul(pixarOnb(hit.n), directionInCosineWeightedHemisphere(blueNoise));
            ray = Ray{p, wi};
            throughput = throughput * albedo;
        }
        else
        {
            const vec3  v = ray.direction;
            const float theta = std::acos(v.y);
            float       c = dot(v, sunDirection);
            c = stdMin(stdMax(c, -1.0f), 1.0f); // clamp(e, low, high) = min(max(e, low), high)
            const float gamma = std::acos(c);
            const vec3  skyRad = {skyRadiance(sky, theta, gamma, 0u), skyRadiance(sky, theta, gamma, 1u), skyRadiance(sky, theta, gamma, 2u)};
            radiance = radiance + throughput * skyRad;
            break;
        }
        bounce += 1u;
    }
    return radiance;
}

// generateCameraRay, wgsl:237-245 (thin lens) and pointInUnitDisk, wgsl:596-600
Ray generateCameraRayWgsl(const float noise[2], const Camera& camera, float u, float v)
{
    const float r = std::sqrt(noise[0]);
    const float theta = 2.0f * PI * noise[1];
    const float lx = camera.lensRadius * (r * std::cos(theta));
    const float ly = camera.lensRadius * (r * std::sin(theta));
    const vec3  lensOffset = lx * load3(camera.right) + ly * load3(camera.up);
    const vec3  origin = load3(camera.origin) + lensOffset;
    const vec3  direction = normalize(load3(camera.lowerLeftCorner) + u * load3(camera.horizontal) + v * load3(camera.vertical) - origin);
    return Ray{origin, direction};
}

// generateCameraRay, common/camera.cpp:44-51 (no lens)
Ray generateCameraRayCpu(const Camera& camera, float u, float v)
{
    const vec3 origin = load3(camera.origin);
    const vec3 direction = load3(camera.lowerLeftCorner) + load3(camera.horizontal) * u + load3(camera.vertical) * v - origin;
    return Ray{origin, normalize(direction)};
}

// animatedBlueNoise, wgsl:603-616
void animatedBlueNoise(const SceneView& scene, std::uint32_t cx, std::uint32_t cy, std::uint32_t frameIdx, std::uint32_t totalSampleCount, float out[2])
{
    const std::uint32_t idx = (cy % scene.bnHeight) * scene.bnWidth + (cx % scene.bnWidth);
    const float         bx = scene.blueNoise[2 * idx], by = scene.blueNoise[2 * idx + 1];
    const std::uint32_t n = frameIdx % totalSampleCount;
    const float         a1 = 0.7548776662466927f;
    const float         a2 = 0.5698402909980532f;
    const float         r2x = fract(a1 * static_cast<float>(n));
    const float         r2y = fract(a2 * static_cast<float>(n));
    out[0] = fract(bx + r2x);
    out[1] = fract(by + r2y);
}

template<typename F>
void parallelRows(int rowBegin, int rowEnd, int numThreads, F&& body)
{
    if (numThreads < 1) numThreads = 1;
    std::atomic<int>         next{rowBegin};
    auto                     work = [&](int tid) {
        for (;;)
        {
            const int row = next.fetch_add(1);
            if (row >= rowEnd) break;
            body(row, tid);
        }
    };
    std::vector<std::thread> threads;
    for (int t = 1; t < numThreads; ++t) threads.emplace_back(work, t);
    work(0);
    for (auto& t : threads) t.join();
}
} // namespace

extern "C" {

int oracle_hardware_concurrency() { return static_cast<int>(std::thread::hardware_concurrency()); }

// createCamera, common/camera.cpp:7-42.  vfov in radians.
void oracle_create_camera(const float* origin, const float* lookAt, float aperture, float focusDistance, float vfovRadians, float aspectRatio, float* out19)
{
    const float halfHeight = focusDistance * std::tan(0.5f * vfovRadians);
    const float halfWidth = aspectRatio * halfHeight;
    const vec3  worldUp = {0.0f, 1.0f, 0.0f};
    const vec3  o = load3(origin);
    const vec3  forward = normalize(load3(lookAt) - o);
    const vec3  right = normalize(cross(forward, worldUp));
    const vec3  up = cross(right, forward);
    const vec3  lowerLeftCorner = o - halfWidth * right - halfHeight * up + focusDistance * forward;
    const vec3  horizontal = 2.0f * halfWidth * right;
    const vec3  vertical = 2.0f * halfHeight * up;
    const vec3  all[6] = {o, lowerLeftCorner, horizontal, vertical, up, right};
    for (int i = 0; i < 6; ++i) out19[3 * i] = all[i].x, out19[3 * i + 1] = all[i].y, out19[3 * i + 2] = all[i].z;
    out19[18] = 0.5f * aperture;
}

int oracle_ray_intersect_aabb(const float* ray6, const float* aabb6, float rayTMax)
{
    const RayAabbIntersector is(Ray{load3(ray6), load3(ray6 + 3)});
    return rayIntersectAabb(is, aabb6, aabb6 + 3, rayTMax) ? 1 : 0;
}

int oracle_ray_intersect_triangle(const float* ray6, const float* tri9, float rayTMax, float* out4)
{
    TriangleHit h{};
    const bool  hit = rayIntersectTriangle(Ray{load3(ray6), load3(ray6 + 3)}, load3(tri9), load3(tri9 + 3), load3(tri9 + 6), rayTMax, h);
    out4[0] = h.p.x, out4[1] = h.p.y, out4[2] = h.p.z, out4[3] = h.t;
    return hit ? 1 : 0;
}

// Batched CPU rayIntersectBvh (ray_intersection.cpp:138-213).  tris = Positions (9 floats).
void oracle_intersect_batch(
    const void* nodes, const float* tris, const float* rays, std::uint64_t numRays, float rayTMax,
    std::uint8_t* outHit, float* outPT, std::uint32_t* outNodes, int numThreads)
{
    SceneView scene{};
    scene.nodes = static_cast<const BvhNode*>(nodes);
    scene.tris = tris;
    scene.triStride = 3;
    const int chunks = static_cast<int>((numRays + 4095) / 4096);
    parallelRows(0, chunks, numThreads, [&](int chunk, int) {
        const std::uint64_t begin = static_cast<std::uint64_t>(chunk) * 4096;
        const std::uint64_t end = begin + 4096 < numRays ? begin + 4096 : numRays;
        for (std::uint64_t i = begin; i < end; ++i)
        {
            float         t = 0.f;
            vec3          p = {0.f, 0.f, 0.f};
            std::uint64_t visited = 0, tested = 0;
            const bool    hit = rayIntersectBvh(scene, Ray{load3(rays + 6 * i), load3(rays + 6 * i + 3)}, rayTMax, nullptr, &t, &p, visited, tested);
            if (outHit) outHit[i] = hit ? 1 : 0;
            if (outPT) outPT[4 * i] = hit ? p.x : 0.f, outPT[4 * i + 1] = hit ? p.y : 0.f, outPT[4 * i + 2] = hit ? p.z : 0.f, outPT[4 * i + 3] = hit ? t : 0.f;
            if (outNodes) outNodes[i] = static_cast<std::uint32_t>(visited);
        }
    });
}

// bvh-visualizer pixel loop (bvh-visualizer/main.cpp:60-78); returns loop wall time in seconds.
double oracle_bvh_visualizer_rows(
    const void* nodes, const float* tris, const float* cam19, int width, int height, int rowBegin, int rowEnd, float rayTMax,
    std::uint32_t* outNodes, int numThreads)
{
    SceneView scene{};
    scene.nodes = static_cast<const BvhNode*>(nodes);
    scene.tris = tris;
    scene.triStride = 3;
    Camera camera;
    std::memcpy(&camera, cam19, sizeof(Camera));
    const auto t0 = std::chrono::steady_clock::now();
    parallelRows(rowBegin, rowEnd, numThreads, [&](int i, int) {
        for (int j = 0; j < width; ++j)
        {
            const float   u = static_cast<float>(j) / static_cast<float>(width);
            const float   v = 1.0f - static_cast<float>(i + 1) / static_cast<float>(height);
            const Ray     ray = generateCameraRayCpu(camera, u, v);
            std::uint64_t visited = 0, tested = 0;
            rayIntersectBvh(scene, ray, rayTMax, nullptr, nullptr, nullptr, visited, tested);
            outNodes[static_cast<std::size_t>(i) * width + j] = static_cast<std::uint32_t>(visited);
        }
    });
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

double oracle_bvh_visualizer(
    const void* nodes, const float* tris, const float* cam19, int width, int height, float rayTMax,
    std::uint32_t* outNodes, int numThreads)
{
    return oracle_bvh_visualizer_rows(nodes, tris, cam19, width, height, 0, height, rayTMax, outNodes, numThreads);
}

float oracle_sky_radiance(const float* skyState40, float theta, float gamma, std::uint32_t channel)
{
    SkyState s;
    std::memcpy(&s, skyState40, sizeof(s));
    return skyRadiance(s, theta, gamma, channel);
}

void oracle_solar_constants(float* out2)
{
    const Constants k = constants();
    out2[0] = k.solarCosThetaMax, out2[1] = k.solarInvPdf;
}

struct OracleScene
{
    const void*          nodes;
    const float*         positionAttributes; // 48 B each
    const void*          vertexAttributes;   // 80 B each
    const std::uint32_t* texDesc;            // (w, h, offset) per texture
    std::uint32_t        numTextures;
    const std::uint32_t* texels;
    std::uint64_t        numTexels;
    const std::uint8_t*  blueNoiseRg8; // 128*128*2
};

struct OracleFrame
{
    std::uint32_t width, height, frameCount, numSamplesPerPixel, numBounces, accumulatedSampleCount;
    std::uint32_t rank, world; // tile ownership as in rf_renderer_set_tile_partition ((tx+ty) % world == rank, 32x32 tiles)
    float         camera[19];
    float         skyState[40];
};

// fsMain, wgsl:34-58 for every (owned) pixel: image (float4 per pixel, the vec3f storage array with
// its 16-byte stride) is zeroed when accumulatedSampleCount == 0 and one sample is added when
// accumulatedSampleCount < numSamplesPerPixel.  counters9: paths, closestRays, shadowRays,
// closestNodes, closestTris, shadowNodes, shadowTris, 0, 0 (accumulated into).
// pathLengths (optional, W*H u8): number of closest-hit rays traced for the pixel.
double oracle_render_frame(const OracleScene* sc, const OracleFrame* fr, float* image, std::uint64_t* counters9, std::uint8_t* pathLengths, int numThreads)
{
    std::vector<float> bn(128 * 128 * 2);
    for (std::size_t i = 0; i < bn.size(); ++i) bn[i] = static_cast<float>(sc->blueNoiseRg8[i]) / 255.0f;
    SceneView scene{};
    scene.nodes = static_cast<const BvhNode*>(sc->nodes);
    scene.tris = sc->positionAttributes;
    scene.triStride = 4;
    scene.vattr = static_cast<const VertexAttributes*>(sc->vertexAttributes);
    scene.texDesc = sc->texDesc;
    scene.numTextures = sc->numTextures;
    scene.texels = sc->texels;
    scene.numTexels = sc->numTexels;
    scene.blueNoise = bn.data();
    scene.bnWidth = 128, scene.bnHeight = 128;
    Camera camera;
    std::memcpy(&camera, fr->camera, sizeof(Camera));
    SkyState sky;
    std::memcpy(&sky, fr->skyState, sizeof(SkyState));

    const std::uint32_t W = fr->width, H = fr->height;
    if (numThreads < 1) numThreads = 1;
    std::vector<Counters> perThread(numThreads);
    const auto            t0 = std::chrono::steady_clock::now();
    parallelRows(0, static_cast<int>(H), numThreads, [&](int py, int tid) {
        Counters& ctr = perThread[tid];
        for (std::uint32_t px = 0; px < W; ++px)
        {
            // vsMain:10-17 + rasteriser: texCoord at the fragment centre.
            const float         u = (static_cast<float>(px) + 0.5f) / static_cast<float>(W);
            const float         v = (static_cast<float>(py) + 0.5f) / static_cast<float>(H);
            const std::uint32_t cx = static_cast<std::uint32_t>(u * static_cast<float>(W));
            const std::uint32_t cy = static_cast<std::uint32_t>(v * static_cast<float>(H));
            const std::uint32_t idx = cy * W + cx;
            float*              px4 = image + 4 * static_cast<std::size_t>(idx);
            if (fr->accumulatedSampleCount == 0u) px4[0] = px4[1] = px4[2] = px4[3] = 0.0f;
            if (((cx / 32u) + (cy / 32u)) % fr->world != fr->rank) continue;
            if (fr->accumulatedSampleCount < fr->numSamplesPerPixel)
            {
                float blueNoise[2];
                animatedBlueNoise(scene, cx, cy, fr->frameCount, fr->numSamplesPerPixel, blueNoise);
                const float jx = blueNoise[0] / static_cast<float>(W), jy = blueNoise[1] / static_cast<float>(H);
                const Ray   primaryRay = generateCameraRayWgsl(blueNoise, camera, u + jx, (1.0f - v) + jy);
                const std::uint64_t before = ctr.closestRays;
                const vec3  c = rayColor(scene, sky, blueNoise, primaryRay, fr->numBounces, ctr);
                ++ctr.paths;
                if (pathLengths) pathLengths[idx] = static_cast<std::uint8_t>(ctr.closestRays - before);
                px4[0] += c.x, px4[1] += c.y, px4[2] += c.z;
            }
        }
    });
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (counters9)
    {
        for (const Counters& c : perThread)
        {
            counters9[0] += c.paths, counters9[1] += c.closestRays, counters9[2] += c.shadowRays, counters9[3] += c.closestNodes;
            counters9[4] += c.closestTris, counters9[5] += c.shadowNodes, counters9[6] += c.shadowTris;
        }
    }
    return seconds;
}

// The rays of one frame: runs fsMain/rayColor (as oracle_render_frame does) for the pixels of every `tileStride`-th 32x32
// tile and logs every ray the path tracer traces for them — the closest-hit ray of each bounce and the shadow ray of each
// hit — as 6 floats (origin, direction) in `outRays` and its kind (0 closest-hit, 1 shadow) in `outKind`, in path order per
// pixel.  Returns the number of rays (at most `capacity` are written).  This is the ray set bench.py hands to the
// reference's own CPU rayIntersectBvh, so that the CPU arm traces the same workload as the GPU arm.
std::uint64_t oracle_frame_rays(
    const OracleScene* sc, const OracleFrame* fr, std::uint32_t tileStride, float* outRays, std::uint8_t* outKind, std::uint64_t capacity, int numThreads)
{
    std::vector<float> bn(128 * 128 * 2);
    for (std::size_t i = 0; i < bn.size(); ++i) bn[i] = static_cast<float>(sc->blueNoiseRg8[i]) / 255.0f;
    SceneView scene{};
    scene.nodes = static_cast<const BvhNode*>(sc->nodes);
    scene.tris = sc->positionAttributes;
    scene.triStride = 4;
    scene.vattr = static_cast<const VertexAttributes*>(sc->vertexAttributes);
    scene.texDesc = sc->texDesc;
    scene.numTextures = sc->numTextures;
    scene.texels = sc->texels;
    scene.numTexels = sc->numTexels;
    scene.blueNoise = bn.data();
    scene.bnWidth = 128, scene.bnHeight = 128;
    Camera camera;
    std::memcpy(&camera, fr->camera, sizeof(Camera));
    SkyState sky;
    std::memcpy(&sky, fr->skyState, sizeof(SkyState));
    const std::uint32_t W = fr->width, H = fr->height, tilesX = (W + 31u) / 32u;
    if (numThreads < 1) numThreads = 1;
    if (tileStride < 1) tileStride = 1;
    // one log per image row, concatenated in row order afterwards: the result does not depend on the thread count
    std::vector<std::vector<float>>        rowRays(H);
    std::vector<std::vector<std::uint8_t>> rowKind(H);
    parallelRows(0, static_cast<int>(H), numThreads, [&](int py, int) {
        Counters ctr;
        ctr.rayLog = &rowRays[py];
        ctr.rayKind = &rowKind[py];
        for (std::uint32_t px = 0; px < W; ++px)
        {
            const float         u = (static_cast<float>(px) + 0.5f) / static_cast<float>(W);
            const float         v = (static_cast<float>(py) + 0.5f) / static_cast<float>(H);
            const std::uint32_t cx = static_cast<std::uint32_t>(u * static_cast<float>(W));
            const std::uint32_t cy = static_cast<std::uint32_t>(v * static_cast<float>(H));
            if (((cy / 32u) * tilesX + (cx / 32u)) % tileStride != 0u) continue;
            float blueNoise[2];
            animatedBlueNoise(scene, cx, cy, fr->frameCount, fr->numSamplesPerPixel, blueNoise);
            const float jx = blueNoise[0] / static_cast<float>(W), jy = blueNoise[1] / static_cast<float>(H);
            const Ray   primaryRay = generateCameraRayWgsl(blueNoise, camera, u + jx, (1.0f - v) + jy);
            rayColor(scene, sky, blueNoise, primaryRay, fr->numBounces, ctr);
        }
    });
    std::uint64_t n = 0;
    for (std::uint32_t y = 0; y < H; ++y)
    {
        const std::uint64_t rays = rowKind[y].size();
        const std::uint64_t take = n + rays <= capacity ? rays : (capacity > n ? capacity - n : 0u);
        if (take && outRays) std::memcpy(outRays + 6 * n, rowRays[y].data(), take * 6 * sizeof(float));
        if (take && outKind) std::memcpy(outKind + n, rowKind[y].data(), take);
        n += rays;
    }
    return n;
}

// ---- the deferred renderer's lighting pass: pt/deferred_renderer_lighting_pass.wgsl ------------------------------
struct OracleDeferred
{
    float         inverseViewReverseZProjection[16]; // column-major
    float         cameraEye[4];
    std::uint32_t width, height, frameCount;
    float         skyState[40];
};

// skyRadiance of the lighting pass (:203-236): the path tracer's plus the solar disk.
static float deferredSkyRadiance(const SkyState& sky, float theta, float gamma, std::uint32_t channel)
{
    const float terrestrialSolarRadius = 0.255f * (PI / 180.0f);
    const float solarDiskRadius = gamma / terrestrialSolarRadius;
    const float solarRadiance = solarDiskRadius <= 1.0f ? sky.solarRadiances[channel] : 0.0f;
    return skyRadiance(sky, theta, gamma, channel) + solarRadiance;
}
static vec3 deferredSky(const SkyState& sky, vec3 v)
{
    const float theta = std::acos(v.y);
    float       c = dot(v, load3(sky.sunDirection));
    c = stdMin(stdMax(c, -1.0f), 1.0f);
    const float gamma = std::acos(c);
    return {deferredSkyRadiance(sky, theta, gamma, 0u), deferredSkyRadiance(sky, theta, gamma, 1u), deferredSkyRadiance(sky, theta, gamma, 2u)};
}
// worldFromUv (:132-138); mat4x4 * vec4 as ((c0 x + c1 y) + c2 z) + c3 w.
static vec3 worldFromUv(const OracleDeferred& un, float uvx, float uvy, float depth)
{
    const float  nx = 2.0f * uvx - 1.0f, ny = 2.0f * (1.0f - uvy) - 1.0f;
    const float* m = un.inverseViewReverseZProjection;
    float        w4[4];
    for (int r = 0; r < 4; ++r) w4[r] = ((m[0 + r] * nx + m[4 + r] * ny) + m[8 + r] * depth) + m[12 + r] * 1.0f;
    return {w4[0] / w4[3], w4[1] / w4[3], w4[2] / w4[3]};
}
// lightSample (:188-200)
static vec3 deferredLightSample(const SceneView& scene, const SkyState& sky, const float u[2], vec3 position, vec3 normal, vec3 albedo, Counters& ctr)
{
    const Constants k = constants();
    const vec3      lightDirection = mul(pixarOnb(load3(sky.sunDirection)), directionInCone(u, k.solarCosThetaMax));
    const vec3      lightIntensity = {sky.solarRadiances[0], sky.solarRadiances[1], sky.solarRadiances[2]};
    const vec3      brdf = albedo * FRAC_1_PI;
    const vec3      reflectance = brdf * dot(normal, lightDirection);
    ++ctr.shadowRays;
    const float lightVisibility = shadowRay(scene, Ray{position, lightDirection}, T_MAX, ctr.shadowNodes, ctr.shadowTris);
    return lightIntensity * reflectance * lightVisibility * k.solarInvPdf;
}

// main + surfaceColor (:96-186) for every pixel.  G-buffer: albedo / encoded normal 4 floats per texel, depth 1 float
// (reverse Z).  sample: 3 floats per pixel (sampleBuffer).  counters9 as oracle_render_frame ("paths" = surface pixels).
double oracle_deferred_lighting(
    const OracleScene* sc, const OracleDeferred* un, const float* albedo4, const float* normal4, const float* depth, float* sample, std::uint64_t* counters9,
    int numThreads)
{
    std::vector<float> bn(128 * 128 * 2);
    for (std::size_t i = 0; i < bn.size(); ++i) bn[i] = static_cast<float>(sc->blueNoiseRg8[i]) / 255.0f;
    SceneView scene{};
    scene.nodes = static_cast<const BvhNode*>(sc->nodes);
    scene.tris = sc->positionAttributes;
    scene.triStride = 4;
    scene.vattr = static_cast<const VertexAttributes*>(sc->vertexAttributes);
    scene.texDesc = sc->texDesc;
    scene.numTextures = sc->numTextures;
    scene.texels = sc->texels;
    scene.numTexels = sc->numTexels;
    scene.blueNoise = bn.data();
    scene.bnWidth = 128, scene.bnHeight = 128;
    scene.deferred = true;
    SkyState sky;
    std::memcpy(&sky, un->skyState, sizeof(SkyState));
    const std::uint32_t W = un->width, H = un->height;
    if (numThreads < 1) numThreads = 1;
    std::vector<Counters> perThread(numThreads);
    const auto            t0 = std::chrono::steady_clock::now();
    parallelRows(0, static_cast<int>(H), numThreads, [&](int py, int tid) {
        Counters& ctr = perThread[tid];
        for (std::uint32_t px = 0; px < W; ++px)
        {
            const std::size_t texel = static_cast<std::size_t>(py) * W + px;
            const float       uvx = (static_cast<float>(px) + 0.5f) / static_cast<float>(W);
            const float       uvy = (static_cast<float>(py) + 0.5f) / static_cast<float>(H);
            const float       depthSample = depth[texel];
            vec3              color = {0.f, 0.f, 0.f};
            if (depthSample == 0.0f)
            {
                const vec3 world = worldFromUv(*un, uvx, uvy, depthSample);
                color = deferredSky(sky, normalize(world - load3(un->cameraEye)));
            }
            else
            {
                const std::uint32_t cx = static_cast<std::uint32_t>(uvx * static_cast<float>(W));
                const std::uint32_t cy = static_cast<std::uint32_t>(uvy * static_cast<float>(H));
                vec3                position = worldFromUv(*un, uvx, uvy, depthSample);
                const float*        en = normal4 + 4 * texel;
                vec3                normal = {2.0f * en[0] - 1.0f, 2.0f * en[1] - 1.0f, 2.0f * en[2] - 1.0f};
                vec3                alb = load3(albedo4 + 4 * texel);
                position = offsetRay(position, normal, true);
                ++ctr.paths;
                // surfaceColor (:142-186), NUM_BOUNCES = 2
                vec3  radiance = {0.f, 0.f, 0.f};
                vec3  throughput = {1.f, 1.f, 1.f};
                float blueNoise[2];
                animatedBlueNoise(scene, cx, cy, un->frameCount, 1u << 20, blueNoise);
                radiance = radiance + throughput * deferredLightSample(scene, sky, blueNoise, position, normal, alb, ctr);
                for (int bounce = 1; bounce < 2; ++bounce)
                {
                    const vec3 wi = mul(pixarOnb(normal), directionInCosineWeightedHemisphere(blueNoise));
                    const Ray  ray{position, wi};
                    throughput = throughput * alb;
                    Intersection hit;
                    ++ctr.closestRays;
                    if (rayIntersectBvh(scene, ray, T_MAX, &hit, nullptr, nullptr, ctr.closestNodes, ctr.closestTris))
                    {
                        position = hit.p;
                        normal = hit.n;
                        alb = textureLookup(scene, hit.textureDescriptorIdx, hit.uv);
                    }
                    else
                    {
                        radiance = radiance + throughput * deferredSky(sky, ray.direction);
                        break;
                    }
                    radiance = radiance + throughput * deferredLightSample(scene, sky, blueNoise, position, normal, alb, ctr);
                }
                color = radiance;
            }
            sample[3 * texel] = color.x, sample[3 * texel + 1] = color.y, sample[3 * texel + 2] = color.z;
        }
    });
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (counters9)
    {
        for (const Counters& c : perThread)
        {
            counters9[0] += c.paths, counters9[1] += c.closestRays, counters9[2] += c.shadowRays, counters9[3] += c.closestNodes;
            counters9[4] += c.closestTris, counters9[5] += c.shadowNodes, counters9[6] += c.shadowTris;
        }
    }
    return seconds;
}

// resolve pass, pt/deferred_renderer_resolve_pass.wgsl:34-50 (the moving average; the tone mapping is oracle_display with
// accumulatedSampleCount 1 on a 4-float copy).
void oracle_deferred_resolve(const float* sample, float* accumulation, std::uint64_t numPixels, std::uint32_t frameCount)
{
    for (std::uint64_t i = 0; i < 3 * numPixels; ++i)
        accumulation[i] = frameCount == 0u ? sample[i] : 0.1f * sample[i] + 0.9f * accumulation[i];
}

// fsMain:59-63 + acesFilmic:278-285, packed as BGRA8 unorm (the reference's swap-chain format).
void oracle_display(const float* image, std::uint64_t numPixels, float accumulatedSampleCount, float exposure, std::uint32_t* outBgra)
{
    for (std::uint64_t i = 0; i < numPixels; ++i)
    {
        std::uint32_t q[3];
        for (int c = 0; c < 3; ++c)
        {
            const float estimator = image[4 * i + c] / accumulatedSampleCount;
            const float x = exposure * estimator;
            const float a = 2.51f, b = 0.03f, cc = 2.43f, d = 0.59f, e = 0.14f;
            float       y = (x * (a * x + b)) / (x * (cc * x + d) + e);
            y = stdMin(stdMax(y, 0.0f), 1.0f); // saturate
            float s = std::pow(y, 1.0f / 2.2f);
            s = (s != s) ? 0.0f : stdMin(stdMax(s, 0.0f), 1.0f);
            q[c] = static_cast<std::uint32_t>(s * 255.0f + 0.5f);
        }
        outBgra[i] = q[2] | (q[1] << 8) | (q[0] << 16) | (255u << 24);
    }
}
}

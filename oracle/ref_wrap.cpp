// TEST INFRASTRUCTURE — not part of the product path.
//
// extern "C" wrapper around the *reference's own* CPU translation units, which oracle/Makefile
// compiles unmodified from where they lie under /root/reference (never copied into this repo):
//   src/common/bvh.cpp              buildBvh                          (bvh.cpp:263-291)
//   src/common/ray_intersection.cpp rayIntersectBvh/Aabb/Triangle     (ray_intersection.cpp:38-213)
//   src/common/camera.cpp           createCamera, generateCameraRay   (camera.cpp:7-51)
//   src/hw-skymodel/hw_skymodel.c   sky_state_new, sky_state_radiance (hw_skymodel.c:141-223)
//   src/pt/blue_noise.c             blueNoiseValues                   (blue_noise.c:3)
// against oracle/shim/glm (glm itself is an un-vendored FetchContent dependency).
// The result, oracle/_ref/libref_oracle.so, pins oracle/oracle.cpp (the restatement) and is the
// CPU baseline ("kind": "reference") timed by bench.py.  Only tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.
#include <common/aabb.hpp>
#include <common/bvh.hpp>
#include <common/camera.hpp>
#include <common/ray.hpp>
#include <common/ray_intersection.hpp>
#include <common/units/angle.hpp>
#include <hw-skymodel/hw_skymodel.h>
#include <pt/blue_noise.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <span>
#include <thread>
#include <vector>

using namespace nlrs;

static_assert(sizeof(BvhNode) == 48);
static_assert(sizeof(Positions) == 36);
static_assert(sizeof(Camera) == 19 * sizeof(float));

namespace
{
Ray toRay(const float* r)
{
    return Ray{.origin = glm::vec3(r[0], r[1], r[2]), .direction = glm::vec3(r[3], r[4], r[5])};
}
} // namespace

extern "C" {

struct RefBvh
{
    Bvh bvh;
};

// buildBvh (bvh.cpp:263).  `tris` = n x Positions (9 floats each).
RefBvh* ref_bvh_build(const float* tris, std::uint64_t n)
{
    std::span<const Positions> span(reinterpret_cast<const Positions*>(tris), n);
    return new RefBvh{buildBvh(span)};
}
std::uint64_t ref_bvh_num_nodes(const RefBvh* b) { return b->bvh.nodes.size(); }
// nodes: numNodes x 48 B; triangleIndices: n x u64 (old index -> new index, bvh.hpp:24-28).
void ref_bvh_copy(const RefBvh* b, void* nodes, std::uint64_t* triangleIndices)
{
    std::memcpy(nodes, b->bvh.nodes.data(), b->bvh.nodes.size() * sizeof(BvhNode));
    for (std::size_t i = 0; i < b->bvh.triangleIndices.size(); ++i)
    {
        triangleIndices[i] = b->bvh.triangleIndices[i];
    }
}
void ref_bvh_free(RefBvh* b) { delete b; }

// createCamera (camera.cpp:7-42); out = the 19 floats of nlrs::Camera (camera.hpp:10-21).
void ref_create_camera(
    const float* origin,
    const float* lookAt,
    float        aperture,
    float        focusDistance,
    float        vfovDegrees,
    float        aspect,
    float*       out19)
{
    const Camera c = createCamera(
        glm::vec3(origin[0], origin[1], origin[2]),
        glm::vec3(lookAt[0], lookAt[1], lookAt[2]),
        aperture,
        focusDistance,
        Angle::degrees(vfovDegrees),
        aspect);
    std::memcpy(out19, &c, sizeof(Camera));
}

// generateCameraRay (camera.cpp:44-51).
void ref_generate_camera_ray(const float* cam19, float u, float v, float* out6)
{
    Camera c;
    std::memcpy(&c, cam19, sizeof(Camera));
    const Ray r = generateCameraRay(c, u, v);
    out6[0] = r.origin.x, out6[1] = r.origin.y, out6[2] = r.origin.z;
    out6[3] = r.direction.x, out6[4] = r.direction.y, out6[5] = r.direction.z;
}

// rayIntersectAabb (ray_intersection.cpp:101-136); aabb6 = min.xyz, max.xyz.
int ref_ray_intersect_aabb(const float* ray6, const float* aabb6, float rayTMax)
{
    const RayAabbIntersector intersector(toRay(ray6));
    Aabb                     box;
    box.min = glm::vec3(aabb6[0], aabb6[1], aabb6[2]);
    box.max = glm::vec3(aabb6[3], aabb6[4], aabb6[5]);
    return rayIntersectAabb(intersector, box, rayTMax) ? 1 : 0;
}

// rayIntersectTriangle (ray_intersection.cpp:38-90); out4 = p.xyz, t.
int ref_ray_intersect_triangle(const float* ray6, const float* tri9, float rayTMax, float* out4)
{
    Positions tri;
    std::memcpy(&tri, tri9, sizeof(Positions));
    Intersection isect{glm::vec3(0.f), 0.f};
    const bool   hit = rayIntersectTriangle(toRay(ray6), tri, rayTMax, isect);
    out4[0] = isect.p.x, out4[1] = isect.p.y, out4[2] = isect.p.z, out4[3] = isect.t;
    return hit ? 1 : 0;
}

// rayIntersectBvh (ray_intersection.cpp:138-213) over a batch of rays, rows split over threads.
// outHit: n x u8; outPT: n x 4 floats (p.xyz, t; untouched values are 0 on miss); outNodes: n x u32.
void ref_intersect_batch(
    const void*    nodes,
    std::uint64_t  numNodes,
    const float*   tris,
    std::uint64_t  numTris,
    const float*   rays,
    std::uint64_t  numRays,
    float          rayTMax,
    std::uint8_t*  outHit,
    float*         outPT,
    std::uint32_t* outNodes,
    int            numThreads)
{
    std::span<const BvhNode>   nodeSpan(static_cast<const BvhNode*>(nodes), numNodes);
    std::span<const Positions> triSpan(reinterpret_cast<const Positions*>(tris), numTris);
    if (numThreads < 1) numThreads = 1;
    std::atomic<std::uint64_t> next{0};
    const std::uint64_t        chunk = 4096;
    auto                       work = [&]() {
        for (;;)
        {
            const std::uint64_t begin = next.fetch_add(chunk);
            if (begin >= numRays) break;
            const std::uint64_t end = begin + chunk < numRays ? begin + chunk : numRays;
            for (std::uint64_t i = begin; i < end; ++i)
            {
                Intersection isect{glm::vec3(0.f), 0.f};
                BvhStats     stats{0};
                const bool   hit =
                    rayIntersectBvh(toRay(rays + 6 * i), nodeSpan, triSpan, rayTMax, isect, &stats);
                if (outHit) outHit[i] = hit ? 1 : 0;
                if (outPT)
                {
                    outPT[4 * i + 0] = hit ? isect.p.x : 0.f;
                    outPT[4 * i + 1] = hit ? isect.p.y : 0.f;
                    outPT[4 * i + 2] = hit ? isect.p.z : 0.f;
                    outPT[4 * i + 3] = hit ? isect.t : 0.f;
                }
                if (outNodes) outNodes[i] = stats.nodesVisited;
            }
        }
    };
    std::vector<std::thread> threads;
    for (int t = 1; t < numThreads; ++t) threads.emplace_back(work);
    work();
    for (auto& t : threads) t.join();
}

// The pixel loop of bvh-visualizer/main.cpp:60-78 (u = j/W, v = 1-(i+1)/H, generateCameraRay,
// rayIntersectBvh with FLT_MAX... here `rayTMax`), rows split over `numThreads` threads.
// Returns the wall time of the loop in seconds (steady_clock around the loop only).
double ref_bvh_visualizer(
    const void*    nodes,
    std::uint64_t  numNodes,
    const float*   tris,
    std::uint64_t  numTris,
    const float*   cam19,
    int            width,
    int            height,
    int            rowBegin,
    int            rowEnd,
    float          rayTMax,
    std::uint32_t* outNodes, // width*height, row-major; only rows [rowBegin,rowEnd) written
    int            numThreads)
{
    std::span<const BvhNode>   nodeSpan(static_cast<const BvhNode*>(nodes), numNodes);
    std::span<const Positions> triSpan(reinterpret_cast<const Positions*>(tris), numTris);
    Camera                     camera;
    std::memcpy(&camera, cam19, sizeof(Camera));
    if (numThreads < 1) numThreads = 1;
    std::atomic<int> nextRow{rowBegin};
    auto             work = [&]() {
        for (;;)
        {
            const int i = nextRow.fetch_add(1);
            if (i >= rowEnd) break;
            for (int j = 0; j < width; ++j)
            {
                const float u = static_cast<float>(j) / static_cast<float>(width);
                const float v = 1.0f - static_cast<float>(i + 1) / static_cast<float>(height);
                const Ray   ray = generateCameraRay(camera, u, v);
                Intersection intersect;
                BvhStats     bvhStats;
                rayIntersectBvh(ray, nodeSpan, triSpan, rayTMax, intersect, &bvhStats);
                outNodes[static_cast<std::size_t>(i) * width + j] = bvhStats.nodesVisited;
            }
        }
    };
    const auto               t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (int t = 1; t < numThreads; ++t) threads.emplace_back(work);
    work();
    for (auto& t : threads) t.join();
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// sky_state_new (hw_skymodel.c:141-180); out33 = params[27], sky_radiances[3], solar_radiances[3].
int ref_sky_state_new(float elevation, float turbidity, const float* albedo3, float* out33)
{
    const sky_params p{
        .elevation = elevation, .turbidity = turbidity, .albedo = {albedo3[0], albedo3[1], albedo3[2]}};
    sky_state s;
    std::memset(&s, 0, sizeof(s));
    const int r = static_cast<int>(sky_state_new(&p, &s));
    std::memcpy(out33, &s, sizeof(s));
    return r;
}
float ref_sky_state_radiance(const float* state33, float theta, float gamma, int ch)
{
    sky_state s;
    std::memcpy(&s, state33, sizeof(s));
    return sky_state_radiance(&s, theta, gamma, static_cast<channel>(ch));
}

const std::uint8_t* ref_blue_noise(std::uint64_t* width, std::uint64_t* height)
{
    *width = blueNoiseWidth;
    *height = blueNoiseHeight;
    return blueNoiseValues;
}

int ref_hardware_concurrency() { return static_cast<int>(std::thread::hardware_concurrency()); }
}

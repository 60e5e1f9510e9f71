// Host-side pieces of the render path that the reference keeps on the CPU: error reporting,
// createCamera, the Hosek-Wilkie sky state, the sampling lookup tables and the binned-SAH BVH
// builder.  No GPU code here; compiled with -ffp-contract=off (strict fp32, see rf_vec.h).
#include "bvh_common.h"
#include "rf_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numbers>

// The two data tables are linked into the library straight from rayfinder_b200/data/*.bin
// (RF_DATA_DIR is passed by the build recipe).
#ifndef RF_DATA_DIR
#error "RF_DATA_DIR must point at rayfinder_b200/data"
#endif
__asm__(".section .rodata\n"
        ".balign 16\n"
        ".global rf_blue_noise_rg8\n"
        "rf_blue_noise_rg8:\n"
        ".incbin \"" RF_DATA_DIR "/blue_noise_128x128_rg8.bin\"\n"
        ".global rf_blue_noise_rg8_end\n"
        "rf_blue_noise_rg8_end:\n"
        ".balign 16\n"
        ".global rf_hw_sky_tables\n"
        "rf_hw_sky_tables:\n"
        ".incbin \"" RF_DATA_DIR "/hw_sky_rgb_tables.bin\"\n"
        ".global rf_hw_sky_tables_end\n"
        "rf_hw_sky_tables_end:\n"
        ".section .text\n");

namespace rfb200
{
namespace
{
thread_local std::string g_lastError;
}

rf_status setError(rf_status code, const char* fmt, ...)
{
    char    buf[1024];
    va_list args;
    va_start(args, fmt);
    std::vsnprintf(buf, sizeof(buf), fmt, args);
    va_end(args);
    g_lastError = buf;
    return code;
}

SolarConstants solarConstants()
{
    // wgsl:68,78-83.  PI / 180f, 0.255f * that, cos(), 2f * PI * (1f - cos) — all f32.
    const float PI = 3.1415927f;
    const float degreesToRadians = PI / 180.0f;
    const float solarRadius = 0.255f * degreesToRadians;
    const float cosThetaMax = std::cos(solarRadius);
    const float invPdf = 2.0f * PI * (1.0f - cosThetaMax);
    return SolarConstants{cosThetaMax, invPdf};
}

namespace
{
// WGSL fract(x) = x - floor(x) (differs from common/math.hpp:7-17 for negative x).
inline float wgslFract(float x) { return x - std::floor(x); }
} // namespace

void buildSampleLutRow(const std::uint32_t n, SampleLutRow& row)
{
    // animatedBlueNoise, wgsl:603-616; blue-noise upload value/255, reference_path_tracer.cpp:174-177.
    const float a1 = 0.7548776662466927f;
    const float a2 = 0.5698402909980532f;
    const float r2x = wgslFract(a1 * static_cast<float>(n));
    const float r2y = wgslFract(a2 * static_cast<float>(n));
    const float twoPi = 2.0f * 3.1415927f;
    for (int b = 0; b < 256; ++b)
    {
        const float bn = static_cast<float>(b) / 255.0f;
        row.ux[b] = wgslFract(bn + r2x);
        row.uy[b] = wgslFract(bn + r2y);
        const float phi = twoPi * row.uy[b];
        row.cosPhi[b] = std::cos(phi);
        row.sinPhi[b] = std::sin(phi);
    }
}

void buildSrgbLut(float out[256])
{
    for (int c = 0; c < 256; ++c)
    {
        out[c] = std::pow(static_cast<float>(c) / 255.0f, 2.2f);
    }
}
} // namespace rfb200

using namespace rfb200;

extern "C" const char* rf_last_error(void) { return g_lastError.c_str(); }

// ---------------------------------------------------------------------------------------------
// createCamera — common/camera.cpp:7-42
// ---------------------------------------------------------------------------------------------
extern "C" rf_status rf_create_camera(
    const float origin[3],
    const float look_at[3],
    const float aperture,
    const float focus_distance,
    const float vfov_radians,
    const float aspect_ratio,
    rf_camera*  out)
{
    if (!origin || !look_at || !out)
    {
        return setError(RF_ERROR_INVALID_ARGUMENT, "rf_create_camera: null argument");
    }
    const float halfHeight = focus_distance * std::tan(0.5f * vfov_radians);
    const float halfWidth = aspect_ratio * halfHeight;

    const V3 eye = v3(origin);
    const V3 worldUp = v3(0.0f, 1.0f, 0.0f);
    const V3 forward = normalize(v3(look_at) - eye);
    const V3 right = normalize(cross(forward, worldUp));
    const V3 up = cross(right, forward);

    // origin - halfWidth*right - halfHeight*up + focusDistance*forward, left to right.
    const V3 lowerLeft = ((eye - halfWidth * right) - halfHeight * up) + focus_distance * forward;
    // 2.0f * halfWidth * right parses as (2.0f * halfWidth) * right.
    const V3 horizontal = (2.0f * halfWidth) * right;
    const V3 vertical = (2.0f * halfHeight) * up;

    const auto store = [](float* dst, V3 v) { dst[0] = v.x, dst[1] = v.y, dst[2] = v.z; };
    store(out->origin, eye);
    store(out->lower_left_corner, lowerLeft);
    store(out->horizontal, horizontal);
    store(out->vertical, vertical);
    store(out->up, up);
    store(out->right, right);
    out->lens_radius = 0.5f * aperture;
    return RF_OK;
}

// ---------------------------------------------------------------------------------------------
// Sky state — hw-skymodel/hw_skymodel.c:19-180 + pt/aligned_sky_state.hpp:43-70
// ---------------------------------------------------------------------------------------------
namespace
{
// Degree-5 Bernstein evaluation of six control values `stride` floats apart (quintic_9/quintic_1,
// hw_skymodel.c:19-63).  Products are formed as ((c * binom) * (1-t)^k) * t^(5-k) and summed in
// control-point order.
float bezier5(const float* ctrl, const std::size_t stride, const float t)
{
    const float t2 = t * t, t3 = t2 * t, t4 = t2 * t2, t5 = t4 * t;
    const float s = 1.0f - t;
    const float s2 = s * s, s3 = s2 * s, s4 = s2 * s2, s5 = s4 * s;
    const float m0 = ctrl[0] * s5;
    const float m1 = ctrl[stride] * 5.0f * s4 * t;
    const float m2 = ctrl[2 * stride] * 10.0f * s3 * t2;
    const float m3 = ctrl[3 * stride] * 10.0f * s2 * t3;
    const float m4 = ctrl[4 * stride] * 5.0f * s * t4;
    const float m5 = ctrl[5 * stride] * t5;
    return m0 + m1 + m2 + m3 + m4 + m5;
}

struct TurbidityLerp
{
    std::size_t lo, hi;
    float       rem;
};
TurbidityLerp turbidityLerp(const float turbidity)
{
    const std::size_t ti = static_cast<std::size_t>(turbidity);
    return TurbidityLerp{ti - 1, ti < 9 ? ti : std::size_t(9), std::fmod(turbidity, 1.0f)};
}

// Bilinear blend over (albedo 0/1) x (turbidity lo/hi) of bezier-interpolated datasets
// (init_params / init_sky_radiance, hw_skymodel.c:65-124).  `block` = floats per turbidity entry.
float blendDataset(
    const float*      data,
    const std::size_t block,
    const std::size_t stride,
    const std::size_t offset,
    const float       turbidity,
    const float       albedo,
    const float       t)
{
    const TurbidityLerp tl = turbidityLerp(turbidity);
    const float*        p0 = data + block * tl.lo + offset;
    const float*        p1 = data + block * tl.hi + offset;
    const float*        p2 = data + block * 10 + block * tl.lo + offset;
    const float*        p3 = data + block * 10 + block * tl.hi + offset;
    const float         s0 = (1.0f - albedo) * (1.0f - tl.rem);
    const float         s1 = (1.0f - albedo) * tl.rem;
    const float         s2 = albedo * (1.0f - tl.rem);
    const float         s3 = albedo * tl.rem;
    float               r = 0.0f;
    r += s0 * bezier5(p0, stride, t);
    r += s1 * bezier5(p1, stride, t);
    r += s2 * bezier5(p2, stride, t);
    r += s3 * bezier5(p3, stride, t);
    return r;
}
} // namespace

extern "C" rf_status rf_sky_state_new(const rf_sky* sky, rf_sky_state* out)
{
    if (!sky || !out)
    {
        return setError(RF_ERROR_INVALID_ARGUMENT, "rf_sky_state_new: null argument");
    }
    std::memset(out, 0, sizeof(*out));

    // AlignedSkyState ctor, aligned_sky_state.hpp:51-63.  Angle::degrees: d * pi_v<float> / 180.0f.
    const float pi = std::numbers::pi_v<float>;
    const float sunZenith = sky->sun_zenith_degrees * pi / 180.0f;
    const float sunAzimuth = sky->sun_azimuth_degrees * pi / 180.0f;
    const V3    sunDir = normalize(v3(
        std::sin(sunZenith) * std::cos(sunAzimuth),
        std::cos(sunZenith),
        -std::sin(sunZenith) * std::sin(sunAzimuth)));
    out->sun_direction[0] = sunDir.x, out->sun_direction[1] = sunDir.y, out->sun_direction[2] = sunDir.z;

    const float elevation = 0.5f * pi - sunZenith;
    const float turbidity = sky->turbidity;

    // Validation, hw_skymodel.c:147-161 (PI = (float)M_PI).
    if (elevation < 0.0f || elevation > pi)
    {
        return setError(RF_ERROR_OUT_OF_RANGE, "sky_state_new: elevation out of range");
    }
    if (turbidity < 1.0f || turbidity > 10.0f)
    {
        return setError(RF_ERROR_OUT_OF_RANGE, "sky_state_new: turbidity out of range");
    }
    for (int c = 0; c < 3; ++c)
    {
        if (sky->albedo[c] < 0.0f || sky->albedo[c] > 1.0f)
        {
            return setError(RF_ERROR_OUT_OF_RANGE, "sky_state_new: albedo out of range");
        }
    }

    const float t = std::pow(elevation / (0.5f * pi), 1.0f / 3.0f);
    constexpr std::size_t PER_CHANNEL = 1080 + 120 + 10;
    for (int c = 0; c < 3; ++c)
    {
        const float* params = rf_hw_sky_tables + PER_CHANNEL * c;
        const float* radiances = params + 1080;
        const float* solar = radiances + 120;
        for (std::size_t i = 0; i < 9; ++i)
        {
            out->params[9 * c + i] = blendDataset(params, 9 * 6, 9, i, turbidity, sky->albedo[c], t);
        }
        out->sky_radiances[c] = blendDataset(radiances, 6, 1, 0, turbidity, sky->albedo[c], t);
        // init_solar_radiance, hw_skymodel.c:126-139.
        const TurbidityLerp tl = turbidityLerp(turbidity);
        out->solar_radiances[c] = solar[tl.lo] * (1.0f - tl.rem) + solar[tl.hi] * tl.rem;
    }
    return RF_OK;
}

// ---------------------------------------------------------------------------------------------
// buildBvh — common/bvh.cpp:81-291 (binned SAH, 12 buckets, depth-first node order)
// ---------------------------------------------------------------------------------------------
namespace
{
struct Prim
{
    Box           box;
    V3            centroid;
    std::uint64_t tri;
};

struct Builder
{
    std::vector<Prim> prims;
    rf_bvh_node*      nodes = nullptr;
    std::uint64_t     numNodes = 0;
    std::uint64_t*    triangleIndices = nullptr;

    static constexpr std::size_t NUM_BUCKETS = BVH_NUM_BUCKETS;

    static std::size_t bucketOf(const Prim& p, int axis, float lo, float hi) { return bvhBucketOf(axisOf(p.centroid, axis), lo, hi); }

    void writeLeaf(std::uint64_t nodeIdx, const Box& box, std::size_t begin, std::size_t end, std::uint64_t firstTri)
    {
        for (std::size_t i = begin; i < end; ++i)
        {
            triangleIndices[prims[i].tri] = firstTri + (i - begin); // bvh.cpp:64-69
        }
        rf_bvh_node& n = nodes[nodeIdx];
        std::memset(&n, 0, sizeof(n));
        n.aabb_min[0] = box.lo.x, n.aabb_min[1] = box.lo.y, n.aabb_min[2] = box.lo.z;
        n.aabb_max[0] = box.hi.x, n.aabb_max[1] = box.hi.y, n.aabb_max[2] = box.hi.z;
        n.triangles_offset = static_cast<std::uint32_t>(firstTri);
        n.second_child_offset = 0;
        n.triangle_count = static_cast<std::uint32_t>(end - begin);
        n.split_axis = 0xFFFFFFFFu; // bvh.cpp:31-42
    }

    // Returns the index of the node built for prims[begin, end); nodes are emitted in pre-order so the
    // first child of an interior node is always nodeIdx + 1 (bvh.cpp:93-94, 236-257).
    std::uint64_t build(std::size_t begin, std::size_t end, std::uint64_t firstTri)
    {
        const std::uint64_t nodeIdx = numNodes++;
        Box                 nodeBox, centroidBox;
        for (std::size_t i = begin; i < end; ++i)
        {
            nodeBox = grow(nodeBox, prims[i].box);
            centroidBox = grow(centroidBox, prims[i].centroid);
        }
        const int         axis = widestAxis(centroidBox);
        const float       cLo = axisOf(centroidBox.lo, axis);
        const float       cHi = axisOf(centroidBox.hi, axis);
        const std::size_t count = end - begin;

        if (area(nodeBox) == 0.0f || cLo == cHi || count == 1)
        {
            writeLeaf(nodeIdx, nodeBox, begin, end, firstTri);
            return nodeIdx;
        }

        const auto  first = prims.begin() + static_cast<std::ptrdiff_t>(begin);
        const auto  last = prims.begin() + static_cast<std::ptrdiff_t>(end);
        std::size_t split = 0;
        if (count < 3)
        {
            split = count / 2; // equal-count split (bvh.cpp:124-137)
            std::nth_element(first, first + static_cast<std::ptrdiff_t>(split), last, [axis](const Prim& a, const Prim& b) {
                return axisOf(a.centroid, axis) < axisOf(b.centroid, axis);
            });
        }
        else
        {
            std::size_t bucketCount[NUM_BUCKETS] = {};
            Box         bucketBox[NUM_BUCKETS];
            for (std::size_t i = begin; i < end; ++i)
            {
                const std::size_t b = bucketOf(prims[i], axis, cLo, cHi);
                bucketCount[b]++;
                bucketBox[b] = grow(bucketBox[b], prims[i].box);
            }

            const int chosen = bvhChooseSplit(bucketCount, bucketBox, nodeBox, count); // the SAH sweep, bvh.cpp:157-214
            if (chosen < 0)
            {
                writeLeaf(nodeIdx, nodeBox, begin, end, firstTri);
                return nodeIdx;
            }
            const std::size_t splitBucket = static_cast<std::size_t>(chosen);
            const auto mid = std::partition(first, last, [=](const Prim& p) {
                return bucketOf(p, axis, cLo, cHi) <= splitBucket;
            });
            split = static_cast<std::size_t>(mid - first);
        }

        build(begin, begin + split, firstTri);
        const std::uint64_t second = build(begin + split, end, firstTri + split);

        rf_bvh_node& n = nodes[nodeIdx];
        std::memset(&n, 0, sizeof(n));
        n.aabb_min[0] = nodeBox.lo.x, n.aabb_min[1] = nodeBox.lo.y, n.aabb_min[2] = nodeBox.lo.z;
        n.aabb_max[0] = nodeBox.hi.x, n.aabb_max[1] = nodeBox.hi.y, n.aabb_max[2] = nodeBox.hi.z;
        n.triangles_offset = 0;
        n.second_child_offset = static_cast<std::uint32_t>(second);
        n.triangle_count = 0;
        n.split_axis = static_cast<std::uint32_t>(axis); // bvh.cpp:44-55
        return nodeIdx;
    }
};
} // namespace

extern "C" rf_status rf_build_bvh(
    const rf_positions* triangles,
    const std::uint64_t num_triangles,
    rf_bvh_node*        out_nodes,
    std::uint64_t*      out_num_nodes,
    std::uint64_t*      out_triangle_indices)
{
    if (!triangles || num_triangles == 0 || !out_nodes || !out_num_nodes || !out_triangle_indices)
    {
        return setError(RF_ERROR_INVALID_ARGUMENT, "rf_build_bvh: null or empty argument");
    }
    if (num_triangles >= std::numeric_limits<std::uint32_t>::max())
    {
        return setError(RF_ERROR_INVALID_ARGUMENT, "rf_build_bvh: too many triangles for u32 offsets");
    }
    Builder b;
    b.prims.reserve(num_triangles);
    for (std::uint64_t i = 0; i < num_triangles; ++i)
    {
        const rf_positions& t = triangles[i];
        const V3            p0 = v3(t.v0), p1 = v3(t.v1), p2 = v3(t.v2);
        // aabb(Positions), aabb.hpp:66-71; centroid = 0.5f * (min + max), aabb.hpp:29.
        const Box box = makeBox(vmin(vmin(p0, p1), p2), vmax(vmax(p0, p1), p2));
        b.prims.push_back(Prim{box, 0.5f * (box.lo + box.hi), i});
    }
    b.nodes = out_nodes;
    b.triangleIndices = out_triangle_indices;
    b.build(0, b.prims.size(), 0);
    *out_num_nodes = b.numNodes;
    return RF_OK;
}

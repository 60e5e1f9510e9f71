// Internal declarations shared by the host translation units and the CUDA translation unit.
#pragma once

#include "../../include/rayfinder_b200.h"
#include "rf_vec.h"

#include <cstdarg>
#include <cstdint>
#include <string>
#include <vector>

namespace rfb200
{
static_assert(sizeof(rf_bvh_node) == 48, "BvhNode layout (common/bvh.hpp:13-21)");
static_assert(sizeof(rf_positions) == 36, "Positions layout (common/triangle_attributes.hpp:7-12)");
static_assert(sizeof(rf_position_attribute) == 48, "PositionAttribute (vertex_attributes.hpp:7-15)");
static_assert(sizeof(rf_vertex_attributes) == 80, "VertexAttributes (vertex_attributes.hpp:17-35)");
static_assert(sizeof(rf_camera) == 76, "Camera (common/camera.hpp:10-21)");
static_assert(sizeof(rf_sky_state) == 160, "AlignedSkyState (pt/aligned_sky_state.hpp:34-41)");

rf_status setError(rf_status code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

// Embedded data blobs (rayfinder_b200/data/*.bin, see tools/extract_reference_data.py).
extern "C" const std::uint8_t rf_blue_noise_rg8[];      // 128*128*2 bytes
extern "C" const std::uint8_t rf_blue_noise_rg8_end[];
extern "C" const float        rf_hw_sky_tables[];       // 3 * (1080 + 120 + 10) floats
extern "C" const float        rf_hw_sky_tables_end[];

constexpr std::uint32_t BLUE_NOISE_WIDTH = 128;
constexpr std::uint32_t BLUE_NOISE_HEIGHT = 128;

// Constants of reference_path_tracer.wgsl:66-83, evaluated once on the host in fp32 exactly as the
// WGSL const-expressions are typed (all operands are f32-suffixed), and passed to the kernels.
struct SolarConstants
{
    float cosThetaMax; // SOLAR_COS_THETA_MAX = cos(0.255f * (PI / 180f))
    float invPdf;      // SOLAR_INV_PDF = 2f * PI * (1f - SOLAR_COS_THETA_MAX)
};
SolarConstants solarConstants();

// Per-sample-index lookup tables of the blue-noise + R2 sequence (wgsl:603-616).  For a fixed
// n = frameCount % numSamplesPerPixel, u = fract(bn + r2(n)) takes 256 values per channel (bn is
// u8/255, reference_path_tracer.cpp:174-177), so every transcendental of the sampling code
// (cos/sin of phi = 2f*PI*u.y; wgsl:572-575, 584-589, 597-599) is tabulated on the host with the
// same libm the oracle uses and the kernels stay transcendental-free on the control-flow path.
struct SampleLutRow
{
    float ux[256];     // fract(bx/255 + fract(a1*n))
    float uy[256];     // fract(by/255 + fract(a2*n))
    float cosPhi[256]; // cos(2f*PI*uy)
    float sinPhi[256]; // sin(2f*PI*uy)
};
void buildSampleLutRow(std::uint32_t n, SampleLutRow& row);
// pow(c/255, 2.2) for c in 0..255 (textureLookup, wgsl:561-563).
void buildSrgbLut(float out[256]);
} // namespace rfb200

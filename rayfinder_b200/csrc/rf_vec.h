// Strict-fp32 3-vector helpers shared by the host code and the CUDA kernels.
//
// Parity contract (DESIGN.md "Arithmetic"): every operation is a single correctly-rounded IEEE fp32
// add/sub/mul/div/sqrt, in the operand order the reference uses (glm 0.9.9.8 scalar formulas for the
// C++ side, left-to-right evaluation for the WGSL side).  The .cu files are compiled with
// -fmad=false and the host files with -ffp-contract=off, so no a*b+c is ever contracted to an FMA.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define RF_HD __host__ __device__ __forceinline__
#else
#define RF_HD inline
#endif

namespace rfb200
{
struct V3
{
    float x, y, z;
};

RF_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
RF_HD V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
RF_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
RF_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
RF_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
RF_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
RF_HD V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }

// glm::dot / WGSL dot: (x + y) + z of the products.
RF_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// glm::cross / WGSL cross.
RF_HD V3 cross(V3 a, V3 b)
{
    return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
// glm::normalize: v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x).
RF_HD V3 normalize(V3 v)
{
#if defined(__CUDA_ARCH__)
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(dot(v, v)));
#else
    const float inv = 1.0f / std::sqrt(dot(v, v));
#endif
    return v * inv;
}
// glm::min / glm::max (std::min / std::max operand order; matters for NaN).
RF_HD float minf(float a, float b) { return (b < a) ? b : a; }
RF_HD float maxf(float a, float b) { return (a < b) ? b : a; }
RF_HD V3    vmin(V3 a, V3 b) { return V3{minf(a.x, b.x), minf(a.y, b.y), minf(a.z, b.z)}; }
RF_HD V3    vmax(V3 a, V3 b) { return V3{maxf(a.x, b.x), maxf(a.y, b.y), maxf(a.z, b.z)}; }
} // namespace rfb200

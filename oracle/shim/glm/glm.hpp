// Minimal stand-in for g-truc/glm 0.9.9.8 (the version the reference pins in
// external/CMakeLists.txt:30-31; glm itself is NOT vendored under /root/reference and is
// absent from this image).  TEST INFRASTRUCTURE ONLY: it exists so that the reference's own
// translation units (src/common/{bvh,ray_intersection,camera}.cpp) compile unmodified into
// oracle/_ref/.  Only the subset those files use is provided; every formula restates the
// published scalar (non-SIMD) glm implementation so that results are bit-identical to a real
// glm build with -ffp-contract=off:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z          (glm/detail/func_geometric.inl compute_dot)
//   cross(x,y)    = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   length(v)     = sqrt(dot(v,v))
//   normalize(v)  = v * inversesqrt(dot(v,v)),  inversesqrt(x) = 1/sqrt(x)
//   min(a,b)      = (b < a) ? b : a ;  max(a,b) = (a < b) ? b : a   (component-wise)
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

namespace glm
{
struct vec2
{
    float x, y;
    vec2() = default;
    constexpr vec2(float xx, float yy) : x(xx), y(yy) {}
    constexpr explicit vec2(float s) : x(s), y(s) {}
    float&       operator[](std::size_t i) { return (&x)[i]; }
    const float& operator[](std::size_t i) const { return (&x)[i]; }
    bool         operator==(const vec2&) const = default;
};

struct vec3
{
    float x, y, z;
    vec3() = default;
    template<typename A, typename B, typename C>
    constexpr vec3(A xx, B yy, C zz)
        : x(static_cast<float>(xx)), y(static_cast<float>(yy)), z(static_cast<float>(zz))
    {
    }
    constexpr explicit vec3(float s) : x(s), y(s), z(s) {}
    float&       operator[](std::size_t i) { return (&x)[i]; }
    const float& operator[](std::size_t i) const { return (&x)[i]; }
    bool         operator==(const vec3&) const = default;
    vec3&        operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};

struct ivec3
{
    int x, y, z;
    ivec3() = default;
    constexpr ivec3(int xx, int yy, int zz) : x(xx), y(yy), z(zz) {}
};

struct vec4
{
    float x, y, z, w;
    vec4() = default;
    constexpr vec4(float xx, float yy, float zz, float ww) : x(xx), y(yy), z(zz), w(ww) {}
    constexpr vec4(const vec3& v, float ww) : x(v.x), y(v.y), z(v.z), w(ww) {}
    bool operator==(const vec4&) const = default;
};

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }

inline float dot(const vec3& a, const vec3& b)
{
    const vec3 t = a * b;
    return t.x + t.y + t.z;
}
inline vec3 cross(const vec3& x, const vec3& y)
{
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec3  normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }

inline float min(float a, float b) { return (b < a) ? b : a; }
inline float max(float a, float b) { return (a < b) ? b : a; }
inline vec3  min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3  max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }

inline float tan(float x) { return std::tan(x); }
} // namespace glm

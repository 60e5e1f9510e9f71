// The deferred renderer's lighting pass as a second integrator on the same traversal kernels:
//   pt/deferred_renderer_lighting_pass.wgsl:96-186   main / worldFromUv / surfaceColor / lightSample
//   pt/deferred_renderer_resolve_pass.wgsl:34-53     exponential moving average of the samples
// The reference's rasteriser fills the G-buffer (albedo, encoded normal, reverse-Z depth); here the G-buffer is
// an INPUT (what textureLoad would return, as floats), so the pass is a drop-in for `lightingPass` + `resolvePass`
// of DeferredRenderer::render (pt/deferred_renderer.cpp:340-375) whoever produced the G-buffer.
//
// One frame = k_deferred_primary -> k_trace(shadow(0) + closest(1)) -> k_shade -> k_trace(shadow(1)) -> k_deferred_resolve.
// It differs from the path tracer (kernels.cuh, FrameParams::deferred) in: the primary hit comes from the G-buffer;
// NUM_BOUNCES = 2; offsetPosition's constants (INT_SCALE 1024, FLOAT_SCALE 1/16384); the solar disk in the sky;
// `radiance += throughput * (lightIntensity * reflectance * visibility * SOLAR_INV_PDF)`; animatedBlueNoise with a
// cycle of 2^20 frames; and the resolve.
#pragma once

#include "kernels.cuh"

namespace rfb200
{
struct DeferredUniforms
{
    float         inverseViewReverseZProjection[16]; // mat4x4f, column-major (WGSL / glm)
    float         cameraEye[4];
    std::uint32_t frameCount;
};

// worldFromUv, deferred_renderer_lighting_pass.wgsl:132-138.  mat4x4 * vec4 = ((c0*x + c1*y) + c2*z) + c3*w.
__device__ __forceinline__ V3 worldFromUv(const DeferredUniforms& un, const float uvx, const float uvy, const float depth)
{
    const float  nx = 2.0f * uvx - 1.0f, ny = 2.0f * (1.0f - uvy) - 1.0f;
    const float* m = un.inverseViewReverseZProjection;
    float        w4[4];
    for (int r = 0; r < 4; ++r) w4[r] = ((m[0 + r] * nx + m[4 + r] * ny) + m[8 + r] * depth) + m[12 + r] * 1.0f;
    return v3(__fdiv_rn(w4[0], w4[3]), __fdiv_rn(w4[1], w4[3]), __fdiv_rn(w4[2], w4[3]));
}

// main, deferred_renderer_lighting_pass.wgsl:96-130, up to the first traversal: sky pixels are finished here, surface
// pixels become queue entries carrying the shadow ray of the primary surface and the bounce ray.
__global__ void __launch_bounds__(BLOCK_THREADS) k_deferred_primary(
    const FrameParams      fp,
    const SceneDevice      scene,
    const DeferredUniforms un,
    const float4* __restrict__ gbufferAlbedo,
    const float4* __restrict__ gbufferNormal,
    const float* __restrict__ gbufferDepth,
    PathQueue              out,
    std::uint32_t*         outCount,
    float4*                radiance,
    unsigned long long*    stats)
{
    const std::uint32_t numPixels = fp.width * fp.height;
    const V3            sunDir = v3(fp.sky.sun_direction);
    std::uint32_t       generated = 0;
    for (std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ((numPixels + 31u) & ~31u); i += gridDim.x * blockDim.x)
    {
        const bool          inside = i < numPixels;
        const std::uint32_t x = inside ? i % fp.width : 0u, y = inside ? i / fp.width : 0u;
        const float         uvx = __fdiv_rn(static_cast<float>(x) + 0.5f, static_cast<float>(fp.width));
        const float         uvy = __fdiv_rn(static_cast<float>(y) + 0.5f, static_cast<float>(fp.height));
        const float         depth = inside ? gbufferDepth[i] : 0.0f;
        const bool          surface = inside && depth != 0.0f;
        if (inside && !surface)
        {
            // reverse-Z depth 0 = nothing rasterised: the sky along the view ray (:103-116)
            const V3 world = worldFromUv(un, uvx, uvy, depth);
            const V3 sky = skyForMiss(fp, normalize(world - v3(un.cameraEye)), sunDir);
            radiance[i] = make_float4(sky.x, sky.y, sky.z, 0.0f);
        }
        const std::uint32_t dst = warpAppend(outCount, surface);
        if (!surface) continue;
        ++generated;
        // coord = vec2u(uv * framebufferSize) indexes the blue noise; it equals the texel (x, y)
        const std::uint32_t cx = static_cast<std::uint32_t>(uvx * static_cast<float>(fp.width));
        const std::uint32_t cy = static_cast<std::uint32_t>(uvy * static_cast<float>(fp.height));
        const std::uint32_t idx = cy * fp.width + cx;
        const V3            position = worldFromUv(un, uvx, uvy, depth);
        const float4        en = gbufferNormal[i], al = gbufferAlbedo[i];
        const V3            n = v3(2.0f * en.x - 1.0f, 2.0f * en.y - 1.0f, 2.0f * en.z - 1.0f);
        const V3            albedo = v3(al.x, al.y, al.z);
        const V3            p = v3(offsetRayComponent(position.x, n.x, true), offsetRayComponent(position.y, n.y, true), offsetRayComponent(position.z, n.z, true));

        // lightSample without the visibility (:188-200) and the bounce direction (:158-160)
        const uchar2        bn = scene.blueNoise[(cy % BLUE_NOISE_HEIGHT) * BLUE_NOISE_WIDTH + (cx % BLUE_NOISE_WIDTH)];
        const SampleLutRow& lut = scene.lut[fp.sampleIndex];
        const float         ux = lut.ux[bn.x];
        const float         cosPhi = lut.cosPhi[bn.y], sinPhi = lut.sinPhi[bn.y];
        const V3            lightDir = sunSampleDirection(fp, scene, idx, sunDir);
        const V3            lightIntensity = v3(fp.sky.solar_radiances[0], fp.sky.solar_radiances[1], fp.sky.solar_radiances[2]);
        const V3            reflectance = (albedo * 0.31830987f) * dot(n, lightDir);
        const V3            contribution = lightIntensity * reflectance;
        const float         hemiSin = __fsqrt_rn(1.0f - ux);
        const V3            wi = onbTransform(n, v3(cosPhi * hemiSin, sinPhi * hemiSin, __fsqrt_rn(ux)));

        out.originPix[dst] = make_float4(p.x, p.y, p.z, __uint_as_float(i)); // sampleBuffer index = texel index (:128)
        out.direction[dst] = make_float4(wi.x, wi.y, wi.z, 0.0f);
        out.throughput[dst] = make_float4(albedo.x, albedo.y, albedo.z, 0.0f);                 // throughput *= albedo (:161)
        out.contribution[dst] = make_float4(contribution.x, contribution.y, contribution.z, 1.0f); // w = 1: throughput was 1
        radiance[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    warpStatAdd(&stats[STAT_PATHS], generated);
}

// resolve pass, deferred_renderer_resolve_pass.wgsl:34-50: accumulation = frameCount == 0 ? sample : 0.1 sample + 0.9 previous.
__global__ void __launch_bounds__(BLOCK_THREADS) k_deferred_resolve(
    const std::uint32_t numPixels, const std::uint32_t frameCount, const float4* __restrict__ sample, float4* accumulation)
{
    for (std::uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += gridDim.x * blockDim.x)
    {
        const float4 cur = sample[i];
        float4       color = cur;
        if (frameCount != 0u)
        {
            const float4 prev = accumulation[i];
            color = make_float4(0.1f * cur.x + 0.9f * prev.x, 0.1f * cur.y + 0.9f * prev.y, 0.1f * cur.z + 0.9f * prev.z, 0.0f);
        }
        accumulation[i] = color;
    }
}
} // namespace rfb200

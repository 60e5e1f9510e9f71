// Host driver of the CUDA render path + its C-ABI: the counterpart of
// nlrs::ReferencePathTracer (pt/reference_path_tracer.{hpp,cpp}) and of the CPU callers of
// nlrs::rayIntersectBvh (bvh-visualizer/main.cpp:60-78, pt/main.cpp:214-218).
//
// Compiled by nvcc for sm_100a only, with -fmad=false (see rf_vec.h).  There is no CPU fallback:
// every entry point fails with RF_ERROR_CUDA when no device is usable.
#include "deferred.cuh"
#include "mega.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

using namespace rfb200;

namespace
{
#define RF_CUDA(expr)                                                                               \
    do                                                                                              \
    {                                                                                               \
        const cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess)                                                                   \
        {                                                                                           \
            return setError(RF_ERROR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
        }                                                                                           \
    } while (0)

template<typename T>
struct DeviceBuffer
{
    T*          ptr = nullptr;
    std::size_t count = 0;

    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { release(); }

    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr, count = 0;
    }
    cudaError_t allocate(std::size_t n)
    {
        release();
        if (n == 0) n = 1;
        const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ptr), n * sizeof(T));
        if (e == cudaSuccess) count = n;
        return e;
    }
};

rf_status selectDevice(int32_t device, int& outDevice, int& numSms)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    {
        cudaGetLastError();
        return setError(RF_ERROR_CUDA, "No CUDA device available: the rayfinder_b200 render path has no CPU fallback.");
    }
    if (device < 0)
    {
        RF_CUDA(cudaGetDevice(&outDevice));
    }
    else
    {
        if (device >= n) return setError(RF_ERROR_INVALID_ARGUMENT, "CUDA device %d out of range (%d devices).", device, n);
        outDevice = device;
    }
    RF_CUDA(cudaSetDevice(outDevice));
    RF_CUDA(cudaDeviceGetAttribute(&numSms, cudaDevAttrMultiProcessorCount, outDevice));
    return RF_OK;
}

// Structural validation of a BVH before it is handed to the kernels: child indices strictly
// increase along any descent (=> termination), leaf ranges lie inside the triangle array, interior
// split axes are 0..2, and the deepest leaf fits the reference's 32-entry stack
// (ray_intersection.cpp:148,194; wgsl:327,375).
rf_status validateBvh(const rf_bvh_node* nodes, std::uint64_t numNodes, std::uint64_t numTriangles, bool& ordered, std::uint32_t* stackEntries = nullptr)
{
    ordered = true;
    if (numNodes == 0 || numNodes >= 0x7FFFFFFFull) return setError(RF_ERROR_INVALID_ARGUMENT, "BVH must have between 1 and 2^31-1 nodes.");
    if (numTriangles >= (1ull << 30)) return setError(RF_ERROR_INVALID_ARGUMENT, "Too many triangles.");
    for (std::uint64_t i = 0; i < numNodes; ++i)
    {
        const rf_bvh_node& n = nodes[i];
        for (int a = 0; a < 3; ++a)
        {
            // finite and min <= max on every axis: the precondition of the NaN-free slab test (traversal.cuh)
            if (!(std::isfinite(n.aabb_min[a]) && std::isfinite(n.aabb_max[a]) && n.aabb_min[a] <= n.aabb_max[a])) ordered = false;
        }
        if (n.triangle_count > 0)
        {
            if (static_cast<std::uint64_t>(n.triangles_offset) + n.triangle_count > numTriangles)
                return setError(RF_ERROR_INVALID_ARGUMENT, "BVH leaf %llu references triangles out of range.", (unsigned long long)i);
            if (n.triangle_count >= (1u << 30))
                return setError(RF_ERROR_INVALID_ARGUMENT, "BVH leaf %llu has too many triangles.", (unsigned long long)i);
        }
        else
        {
            if (n.split_axis > 2u) return setError(RF_ERROR_INVALID_ARGUMENT, "BVH interior node %llu has split axis %u.", (unsigned long long)i, n.split_axis);
            if (i + 1 >= numNodes || n.second_child_offset <= i + 1 || n.second_child_offset >= numNodes)
                return setError(RF_ERROR_INVALID_ARGUMENT, "BVH interior node %llu has child indices out of order.", (unsigned long long)i);
        }
    }
    // Depth check (explicit stack).  Every node must have exactly one parent: a node reached twice means the records
    // describe a DAG, whose root-to-leaf paths can be exponentially many — rejected at the second visit, which keeps
    // this walk O(numNodes).
    std::vector<std::pair<std::uint32_t, std::uint32_t>> stack;
    std::vector<std::uint8_t>                            reached(numNodes, 0);
    stack.emplace_back(0u, 0u);
    reached[0] = 1;
    std::uint32_t maxPending = 0;
    while (!stack.empty())
    {
        auto [idx, pending] = stack.back();
        stack.pop_back();
        const rf_bvh_node& n = nodes[idx];
        if (n.triangle_count == 0)
        {
            // descending into one child leaves the other pending on the traversal stack
            maxPending = std::max(maxPending, pending + 1);
            if (maxPending > static_cast<std::uint32_t>(RF_STACK_SIZE)) break;
            for (const std::uint32_t child : {idx + 1u, n.second_child_offset})
            {
                if (reached[child])
                    return setError(RF_ERROR_INVALID_ARGUMENT, "BVH node %u has more than one parent.", child);
                reached[child] = 1;
                stack.emplace_back(child, pending + 1);
            }
        }
    }
    if (maxPending > static_cast<std::uint32_t>(RF_STACK_SIZE))
        return setError(RF_ERROR_INVALID_ARGUMENT, "BVH depth %u exceeds the traversal stack of %d entries.", maxPending, RF_STACK_SIZE);
    if (stackEntries) *stackEntries = std::max(maxPending, 1u);
    return RF_OK;
}

// Scheduling knobs of the persistent traversal loop (rf_renderer_set_tuning / rf_renderer_set_option change them).
TraceTuning defaultTuning() { return TraceTuning{4u, 4u, 16u, 0u, 1u, 0u}; }

bool sameParams(const rf_render_parameters& a, const rf_render_parameters& b)
{
    // RenderParameters::operator== (reference_path_tracer.hpp:42): member-wise, floats by value.
    const auto eq3 = [](const float* x, const float* y) { return x[0] == y[0] && x[1] == y[1] && x[2] == y[2]; };
    return a.framebuffer_width == b.framebuffer_width && a.framebuffer_height == b.framebuffer_height &&
           eq3(a.camera.origin, b.camera.origin) && eq3(a.camera.lower_left_corner, b.camera.lower_left_corner) &&
           eq3(a.camera.horizontal, b.camera.horizontal) && eq3(a.camera.vertical, b.camera.vertical) &&
           eq3(a.camera.up, b.camera.up) && eq3(a.camera.right, b.camera.right) &&
           a.camera.lens_radius == b.camera.lens_radius &&
           a.sampling_params.num_samples_per_pixel == b.sampling_params.num_samples_per_pixel &&
           a.sampling_params.num_bounces == b.sampling_params.num_bounces && a.sky.turbidity == b.sky.turbidity &&
           eq3(a.sky.albedo, b.sky.albedo) && a.sky.sun_zenith_degrees == b.sky.sun_zenith_degrees &&
           a.sky.sun_azimuth_degrees == b.sky.sun_azimuth_degrees && a.exposure == b.exposure;
}

rf_status validateParams(const rf_render_parameters& p, std::uint32_t maxW, std::uint32_t maxH)
{
    if (p.framebuffer_width == 0 || p.framebuffer_height == 0 || p.framebuffer_width > maxW || p.framebuffer_height > maxH)
        return setError(RF_ERROR_INVALID_ARGUMENT, "Framebuffer size %ux%u outside (0, %ux%u].", p.framebuffer_width, p.framebuffer_height, maxW, maxH);
    if (p.sampling_params.num_samples_per_pixel == 0 || p.sampling_params.num_samples_per_pixel > 65536u)
        return setError(RF_ERROR_INVALID_ARGUMENT, "numSamplesPerPixel must be in [1, 65536].");
    if (p.sampling_params.num_bounces == 0 || p.sampling_params.num_bounces > 1024u)
        return setError(RF_ERROR_INVALID_ARGUMENT, "numBounces must be in [1, 1024].");
    return RF_OK;
}
} // namespace

// Compile-time scheduling variants of the traversal kernel, selected per launch (tuning only).
template<int V, int BLOCK, int STACK = RF_STACK_SIZE>
void launchTraceV(int grid, cudaStream_t s, const FrameParams& fp, const SceneDevice& scene, const PathQueue& closestQueue,
                  const std::uint32_t* closestCount, HitRecord* hits, const PathQueue& shadowQueue, const std::uint32_t* shadowCount,
                  float4* radiance, std::uint32_t* cursor, const StragglerBuffer& stragglers, unsigned long long* stats)
{
    k_trace<V, BLOCK, STACK><<<grid, BLOCK, 0, s>>>(fp, scene, closestQueue, closestCount, hits, shadowQueue, shadowCount, radiance, cursor, stragglers, stats);
}
// `stackEntries` = what the scene needs (validateBvh).  The default kernel (variant 3, 256 threads) exists with 23-,
// 31- and 32-entry stacks: the smaller ones leave more of the SM's 256 KB to L1 (traversal.cuh, traceRays).
template<typename... Args>
void launchTrace(int variant, int block, std::uint32_t stackEntries, Args&&... args)
{
    if (block == 256 && (variant & 15) == TRACE_DEFAULT_VARIANT && stackEntries <= 31u)
    {
        if (stackEntries <= 23u)
            launchTraceV<TRACE_DEFAULT_VARIANT, 256, 23>(args...);
        else
            launchTraceV<TRACE_DEFAULT_VARIANT, 256, 31>(args...);
        return;
    }
    if (block == 64)
    {
        if ((variant & 15) == 2) launchTraceV<2, 64>(args...);
        else launchTraceV<TRACE_DEFAULT_VARIANT, 64>(args...);
    }
    else if (block == 128)
    {
        if ((variant & 15) == 2) launchTraceV<2, 128>(args...);
        else launchTraceV<TRACE_DEFAULT_VARIANT, 128>(args...);
    }
    else
    {
        switch (variant & 15)
        {
        case 1: launchTraceV<1, 256>(args...); break;
        case 2: launchTraceV<2, 256>(args...); break;
        case 11: launchTraceV<11, 256>(args...); break;
        default: launchTraceV<TRACE_DEFAULT_VARIANT, 256>(args...); break;
        }
    }
}

template<typename... Args>
void launchTracePairs(int variant, int grid, cudaStream_t s, Args&&... args)
{
    switch (variant & 7)
    {
    case 1: k_trace_pairs<1, 256><<<grid, 256, 0, s>>>(args...); break;
    case 5: k_trace_pairs<5, 256><<<grid, 256, 0, s>>>(args...); break;
    case 7: k_trace_pairs<7, 256><<<grid, 256, 0, s>>>(args...); break;
    default: k_trace_pairs<3, 256><<<grid, 256, 0, s>>>(args...); break;
    }
}

// The persistent kernel (mega.cuh) with the stack the scene needs, like launchTrace; blocks of 256 threads (4 per SM, 7
// traversal warps + 1 shading warp each) or 512 threads (2 per SM, 15 + 1).
template<int V, int BLOCK, int STACK, typename... Args>
void launchMegaV(int grid, cudaStream_t s, Args&&... args)
{
    constexpr std::size_t bytes = megaSharedBytes<BLOCK, STACK>();
    static bool           configured = false; // (per instantiation; the attribute is per device function, set again after a device change is harmless)
    if (!configured)
    {
        cudaFuncSetAttribute(k_mega<V, BLOCK, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
        configured = true;
    }
    k_mega<V, BLOCK, STACK><<<grid, BLOCK, bytes, s>>>(args...);
}
template<typename... Args>
void launchMega(int variant, int block, std::uint32_t stackEntries, int grid, cudaStream_t s, Args&&... args)
{
    if (block == 512)
    {
        if (stackEntries <= 23u)
            launchMegaV<TRACE_DEFAULT_VARIANT, 512, 23>(grid, s, args...);
        else if (stackEntries <= 31u)
            launchMegaV<TRACE_DEFAULT_VARIANT, 512, 31>(grid, s, args...);
        else
            launchMegaV<TRACE_DEFAULT_VARIANT, 512, RF_STACK_SIZE>(grid, s, args...);
    }
    else if ((variant & 15) == 2)
        launchMegaV<2, 256, RF_STACK_SIZE>(grid, s, args...);
    else if ((variant & 15) == 1)
        launchMegaV<1, 256, RF_STACK_SIZE>(grid, s, args...);
    else if (stackEntries <= 23u)
        launchMegaV<TRACE_DEFAULT_VARIANT, 256, 23>(grid, s, args...);
    else if (stackEntries <= 31u)
        launchMegaV<TRACE_DEFAULT_VARIANT, 256, 31>(grid, s, args...);
    else
        launchMegaV<TRACE_DEFAULT_VARIANT, 256, RF_STACK_SIZE>(grid, s, args...);
}

// =================================================================================================
struct rf_renderer
{
    int          device = 0;
    int          numSms = 0;
    cudaStream_t stream = nullptr; // default stream unless rf_renderer_set_stream is called

    // scene
    DeviceBuffer<PackedNode>    nodes;
    DeviceBuffer<float4>        tris, vattr;
    bool                        ordered = true;
    TraceTuning                 tuning = defaultTuning();
    DeviceBuffer<PairRecord>    pairRecords;  // child-pair records (pair_records.h); empty when the scene's leaves do not fit a link
    PairSceneDevice             pairsDev{};
    // 0 / 1: one node per visit (traversal.cuh, the default: measured faster, DESIGN.md "Child-pair records"), 2: child-pair
    // records (traversal_pairs.cuh) when the scene has them
    int                         traceKernel = 0;
    int                         pairVariant = PAIR_DEFAULT_VARIANT;
    bool                        usePairs() const { return pairsDev.records != nullptr && traceKernel == 2; }
    DeviceBuffer<uint4>         texDesc;
    DeviceBuffer<std::uint32_t> texels;
    DeviceBuffer<uchar2>        blueNoise;
    DeviceBuffer<SampleLutRow>  lut;
    DeviceBuffer<float>         srgbLut;
    std::uint32_t               numTextures = 0;
    std::uint64_t               numTexels = 0;

    // frame state
    std::uint32_t               maxW = 0, maxH = 0;
    DeviceBuffer<float4>        image, radiance;
    DeviceBuffer<unsigned long long> stats;
    DeviceBuffer<std::uint32_t> display;

    // A frame is traced as `numSubFrames` independent sub-frames (the rank's tiles dealt round-robin), each
    // with its own queues and stream: while one sub-frame's traversal launch drains its longest rays (a
    // latency-bound tail of ~0.25 ms per launch), the other sub-frame's kernels fill the machine.
    struct SubFrame
    {
        cudaStream_t                stream = nullptr; // sub-frame 0 runs on the renderer's stream
        cudaEvent_t                 done = nullptr;
        DeviceBuffer<float4>        queueMem; // 2 queues x 4 arrays x capacity
        DeviceBuffer<HitRecord>     hits;
        DeviceBuffer<std::uint32_t> ownedTiles;
        DeviceBuffer<std::uint32_t> counters; // see counterSlots()
        DeviceBuffer<StragglerRecord> stragglers; // rays handed over by the tails of the traversal launches (traversal.cuh)
        DeviceBuffer<float4>        pathRecords; // persistent-kernel mode: blocks x slots path records (mega.cuh), allocated on first use
        std::uint64_t               pathRecordSlots = 0;
        PathQueue                   queues[2]{};
        std::uint64_t               capacity = 0; // paths
        std::uint32_t               numOwnedTiles = 0;
    };
    static constexpr int MAX_SUBFRAMES = 4;
    SubFrame    sub[MAX_SUBFRAMES];
    // The deferred renderer's lighting pass (deferred.cuh): G-buffer copies, its own queues and EMA buffer; allocated
    // by the first rf_renderer_render_deferred_lighting call.
    struct Deferred
    {
        DeviceBuffer<float4>        albedo, normal, queueMem, accumulation;
        DeviceBuffer<float>         depth;
        DeviceBuffer<HitRecord>     hits;
        DeviceBuffer<std::uint32_t> counters;
        DeviceBuffer<SampleLutRow>  lutRow;
        std::uint64_t               capacity = 0;
        std::uint32_t               width = 0, height = 0;
        float                       exposure = 1.0f;
    } deferred;
    // Multi-GPU exchange over peer memory (rf_renderer_hdr_ipc_handle / rf_renderer_set_hdr_peer).  The root owns
    // `exchange`: TWO full frames.  Frame N of every rank (the root included) stores its owned pixels into half N & 1, so
    // the root can read frame N while the other ranks already write frame N + 1 into the other half; they reach half
    // N & 1 again only after the frame barrier of N + 1, which the root enters after its read (stream order).
    DeviceBuffer<float4> exchange;
    float4*       peerImage = nullptr;  // both halves: the root's own `exchange.ptr`, or its mapping through CUDA IPC on the other ranks
    bool          exportedRoot = false; // this renderer owns the exchange buffer
    std::uint32_t exchangeEpoch = 0;    // frames accumulated since the exchange was set up (same on every rank)
    std::uint64_t exchangeStride() const { return static_cast<std::uint64_t>(maxW) * maxH; }
    float4*       exchangeTarget() const { return peerImage ? peerImage + (exchangeEpoch & 1u) * exchangeStride() : nullptr; }
    // what read_hdr / read_display show: the last complete exchanged frame on the root of an exchange, else the local image
    const float4* presentedImage() const
    {
        return exportedRoot && exchangeEpoch > 0 ? exchange.ptr + ((exchangeEpoch - 1u) & 1u) * exchangeStride() : image.ptr;
    }
    std::uint64_t kernelLaunches = 0;   // kernels launched by render() since the last reset_stats
    int         numSubFrames = 2;       // in effect (updateTiles)
    int         requestedSubFrames = 0; // 0: automatic
    // The frame as one persistent kernel with block-local path loops (mega.cuh): 0 never, 1 always, 2 automatic = when this GPU
    // owns at most ~1.2 M pixels (a 1080p frame split over 2-8 GPUs), where the ends of the staged pipeline's nine traversal
    // launches weigh most (measured on one B200, persistent vs staged: 672x384 2.42 vs 3.18 ms, 960x540 4.37 vs 5.14 ms,
    // 1360x768 8.49 vs 8.72 ms; the full 1080p frame 16.5 vs 15.9 ms).
    int         megaMode = 2;
    cudaEvent_t forkEvent = nullptr;

    rf_render_parameters params{};
    rf_sky_state         skyState{};
    std::uint32_t        lutRows = 0;
    std::uint32_t        frameCount = 0;
    std::uint32_t        accumulated = 0;
    std::uint32_t        rank = 0, world = 1;
    std::uint32_t        tilesX = 0;
    bool                 tilesDirty = true;

    // timing
    struct Timed
    {
        cudaEvent_t              begin = nullptr, end = nullptr;
        std::vector<cudaEvent_t> stages; // recorded between stages when stage timing is on
        std::uint32_t            stagesUsed = 0, bounces = 0;
    };
    std::vector<Timed>  eventPool;
    std::deque<Timed>   pending;
    std::deque<float>   durationsMs; // last <= 30 completed passes
    double              totalMs = 0.0;
    std::uint64_t       frames = 0;
    bool                stageTiming = false;
    double              msTrace = 0, msShade = 0, msOther = 0;

    ~rf_renderer()
    {
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        const auto destroy = [](Timed& t) {
            cudaEventDestroy(t.begin);
            cudaEventDestroy(t.end);
            for (auto e : t.stages) cudaEventDestroy(e);
        };
        for (auto& t : eventPool) destroy(t);
        for (auto& t : pending) destroy(t);
        for (auto& sf : sub)
        {
            if (sf.stream) cudaStreamDestroy(sf.stream);
            if (sf.done) cudaEventDestroy(sf.done);
        }
        if (forkEvent) cudaEventDestroy(forkEvent);
        if (peerImage && !exportedRoot) cudaIpcCloseMemHandle(peerImage);
    }

    void drainTimings(bool wait)
    {
        while (!pending.empty())
        {
            Timed& t = pending.front();
            if (wait)
            {
                cudaEventSynchronize(t.end);
            }
            else if (cudaEventQuery(t.end) != cudaSuccess)
            {
                cudaGetLastError();
                break;
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, t.begin, t.end);
            durationsMs.push_back(ms);
            if (durationsMs.size() > 30) durationsMs.pop_front(); // reference_path_tracer.cpp:689-693
            totalMs += ms;
            if (t.stagesUsed == 4 + 2 * t.bounces)
            {
                // stage events: [0] before raygen, [1] after raygen, then per bounce (after k_trace, after
                // k_shade), then after the last k_trace, after k_accumulate.
                const auto span = [&](std::uint32_t a, std::uint32_t b) {
                    float x = 0.f;
                    cudaEventElapsedTime(&x, t.stages[a], t.stages[b]);
                    return static_cast<double>(x);
                };
                msOther += span(0, 1);
                std::uint32_t e = 1;
                for (std::uint32_t b = 0; b < t.bounces; ++b, e += 2)
                {
                    msTrace += span(e, e + 1);
                    msShade += span(e + 1, e + 2);
                }
                msTrace += span(e, e + 1);
                msOther += span(e + 1, e + 2);
                if (stageDebug)
                {
                    // per-launch spans of one frame: raygen, then (trace, shade) per bounce, last trace, accumulate
                    std::fprintf(stderr, "[stages] total %.3f:", ms);
                    for (std::uint32_t k = 0; k + 1 < t.stagesUsed; ++k) std::fprintf(stderr, " %.3f", span(k, k + 1));
                    std::fprintf(stderr, "\n");
                }
            }
            t.stagesUsed = 0;
            eventPool.push_back(std::move(t));
            pending.pop_front();
        }
    }

    rf_status applyParams(const rf_render_parameters& p)
    {
        // the sky state first: a rejected sky must leave the renderer as it was (params, sky state and accumulation)
        rf_sky_state    newSky{};
        const rf_status st = rf_sky_state_new(&p.sky, &newSky);
        if (st != RF_OK) return st;
        skyState = newSky;
        if (p.framebuffer_width != params.framebuffer_width || p.framebuffer_height != params.framebuffer_height ||
            p.sampling_params.num_bounces != params.sampling_params.num_bounces)
            tilesDirty = true; // (the automatic schedule depends on both)
        params = p;
        accumulated = 0; // reset the temporal accumulation (reference_path_tracer.cpp:561)
        // Sampling tables for every n = frameCount % numSamplesPerPixel.
        const std::uint32_t spp = p.sampling_params.num_samples_per_pixel;
        if (spp != lutRows)
        {
            std::vector<SampleLutRow> rows(spp);
            for (std::uint32_t n = 0; n < spp; ++n) buildSampleLutRow(n, rows[n]);
            RF_CUDA(cudaStreamSynchronize(stream));
            RF_CUDA(lut.allocate(spp));
            RF_CUDA(cudaMemcpy(lut.ptr, rows.data(), spp * sizeof(SampleLutRow), cudaMemcpyHostToDevice));
            lutRows = spp;
        }
        return RF_OK;
    }

    rf_status updateTiles()
    {
        if (!tilesDirty) return RF_OK;
        tilesX = (params.framebuffer_width + TILE - 1) / TILE;
        const std::uint32_t tilesY = (params.framebuffer_height + TILE - 1) / TILE;
        std::vector<std::uint32_t> owned[MAX_SUBFRAMES];
        std::uint32_t              k = 0;
        ownedTileCount = 0;
        for (std::uint32_t ty = 0; ty < tilesY; ++ty)
            for (std::uint32_t tx = 0; tx < tilesX; ++tx)
                if ((tx + ty) % world == rank) ++ownedTileCount;
        // Automatic number of tile sets: 1 for the persistent kernel; for the staged pipeline 2 (one set's launch tails overlap the
        // other's work) — except between ~1.2 M and ~3 M owned pixels while warps walk their last rays in place (tuning.walkInPlace):
        // the tails are short then, and one set of full-size launches measured faster at 1920x1080 (15.86 vs 15.94 ms; at
        // 1024^2 and 3840x2160 two sets stay ahead by ~1 %).
        const std::uint64_t ownedPixels = static_cast<std::uint64_t>(ownedTileCount) * TILE_PIXELS;
        const bool          oneStagedSet = tuning.walkInPlace != 0u && ownedPixels > 1200000ull && ownedPixels <= 3000000ull;
        numSubFrames = requestedSubFrames > 0 ? requestedSubFrames : (useMega() || oneStagedSet ? 1 : 2);
        for (std::uint32_t ty = 0; ty < tilesY; ++ty)
            for (std::uint32_t tx = 0; tx < tilesX; ++tx)
                if ((tx + ty) % world == rank) owned[k++ % static_cast<std::uint32_t>(numSubFrames)].push_back(ty * tilesX + tx);
        RF_CUDA(cudaStreamSynchronize(stream));
        for (int i = 0; i < MAX_SUBFRAMES; ++i)
        {
            SubFrame& sf = sub[i];
            if (sf.stream) RF_CUDA(cudaStreamSynchronize(sf.stream));
            sf.numOwnedTiles = i < numSubFrames ? static_cast<std::uint32_t>(owned[i].size()) : 0u;
            if (sf.numOwnedTiles == 0) continue;
            const std::uint64_t need = static_cast<std::uint64_t>(sf.numOwnedTiles) * TILE_PIXELS;
            if (need > sf.capacity)
            {
                RF_CUDA(sf.queueMem.allocate(need * 8));
                RF_CUDA(sf.hits.allocate(need));
                RF_CUDA(sf.ownedTiles.allocate(sf.numOwnedTiles));
                if (!sf.counters.ptr) RF_CUDA(sf.counters.allocate(counterSlots(1024)));
                if (!sf.stragglers.ptr) RF_CUDA(sf.stragglers.allocate(stragglerCapacity()));
                for (int q = 0; q < 2; ++q)
                {
                    float4* base = sf.queueMem.ptr + static_cast<std::uint64_t>(q) * 4 * need;
                    sf.queues[q] = PathQueue{base, base + need, base + 2 * need, base + 3 * need};
                }
                sf.capacity = need;
            }
            RF_CUDA(cudaMemcpy(sf.ownedTiles.ptr, owned[i].data(), owned[i].size() * sizeof(std::uint32_t), cudaMemcpyHostToDevice));
        }
        tilesDirty = false;
        return RF_OK;
    }

    int variant = TRACE_DEFAULT_VARIANT; // scheduling variant of the traversal kernels (traversal.cuh)
    int traceBlocksPerSm = 0; // persistent traversal blocks per SM in units of 256 threads (0: automatic; 4 fill an SM: <= 64 registers, 32 KB stack)
    // Automatic: a launch of a full frame fills the machine (4); the launches of a small frame's two tile sets take half of it
    // each (2), so that both sets' kernels are resident together and one set's tail overlaps the other's steady state
    // (measured at 672x384 / 960x540, the shares of 8 / 4 GPUs: 3.15 / 5.04 ms against 3.27 / 5.24 ms with one set on 4).
    int effectiveBlocksPerSm() const { return traceBlocksPerSm > 0 ? traceBlocksPerSm : (smallFrame() && numSubFrames == 2 ? 2 : 4); }
    int traceBlock = 256;     // threads per traversal block (64, 128 or 256)
    int gridFor(int blocksPerSm) const { return numSms * blocksPerSm; }
    // Tail policy of the traversal launches: warps left with <= evictMax rays once the queue is dry hand them to a
    // small follow-up launch (0 = off).  On by default when the frame runs as several tile sets, where the SMs a
    // tail frees are used by the other sets.
    // Automatic scheduling, from measurements on B200 (Sponza, 8 bounces; DESIGN.md "Tails"): a frame is always traced as
    // two tile sets on two streams.  With more than ~0.6 M paths per GPU (the 1080p frame on 1-2 GPUs) every traversal launch
    // fills the machine and ends its rays in place; below that (a 1080p frame split over 4-8 GPUs) a launch is mostly tail:
    // each set's launches take half the SM slots, so both sets are resident together, and hand their tails over.
    std::uint32_t stackEntries = RF_STACK_SIZE; // deepest traversal stack the scene can produce (validateBvh)
    // option "trace_stack" = 32 forces the reference-sized stack (A/B runs)
    std::uint32_t forcedStackEntries = 0;
    std::uint32_t traceStackEntries() const { return std::max(forcedStackEntries, stackEntries); }
    std::uint32_t evictDelay = 4; // rounds a warp keeps its last rays before handing them over (measured: 4 lets the many short ones end in place)
    bool          stageDebug = false;
    std::uint32_t megaSlots = 0; // path slots per block of the persistent kernel (0: automatic; option "mega_slots")
    // threads per block of the persistent kernel (option "mega_block"): 512 = 2 blocks per SM with 15 traversal warps + 1 shading
    // warp each (measured 2.42 / 4.72 / 16.5 ms at 672x384 / 960x540 / 1080p); 256 = 4 per SM with 7 + 1 (2.47 / 4.92 / 17.6 ms)
    int           megaBlock = 512;
    int           stragglerWindowMode = STRAGGLER_DIRECT; // how the tail kernel fetches its node windows (straggler.cuh)
    int           evictMax = -1; // -1: automatic
    std::uint32_t ownedTileCount = 0;
    bool          smallFrame() const { return static_cast<std::uint64_t>(ownedTileCount) * TILE_PIXELS <= 600000ull; }
    // Automatic: the persistent kernel pays where a frame is many short launches — few pixels per GPU AND several bounces.  With one
    // or two bounces the staged pipeline has only two or three launches, the first of them over coherent primary rays, and stays
    // ahead (1024^2: 1.00 vs 1.17 ms at 1 bounce, 4.64 vs 4.76 ms at 4, 8.83 vs 8.65 ms at 8; 512^2: 0.87 vs 0.85 ms at 2 bounces,
    // 1.76 vs 1.43 ms at 4; profiles/r02_sweep_1gpu.json against r01b_sweep_1gpu.json).
    bool          useMega() const
    {
        if (megaMode != 2) return megaMode == 1;
        const std::uint64_t pixels = static_cast<std::uint64_t>(ownedTileCount) * TILE_PIXELS;
        const std::uint32_t bounces = params.sampling_params.num_bounces;
        return pixels <= 1200000ull && (bounces >= 6u || (bounces >= 3u && pixels > 150000ull && pixels <= 600000ull));
    }
    std::uint32_t stragglerCapacity() const { return static_cast<std::uint32_t>(numSms) * 64u * 8u; }
    // (the tail hand-over belongs to the per-node kernel; the pair kernel ends every ray on its lane)
    std::uint32_t effectiveEvictMax() const
    {
        if (usePairs() && !useMega()) return 0u;
        // automatic: off while warps walk their last ray in place (tuning.walkInPlace, the default: measured 1200 vs 1152 Mrays/s at
        // 672x384 with the hand-over on top); without it, 8 for a small frame
        return evictMax >= 0 ? static_cast<std::uint32_t>(evictMax) : (smallFrame() && tuning.walkInPlace == 0u ? 8u : 0u);
    }
};

extern "C" rf_status rf_renderer_create(
    const rf_renderer_descriptor* desc,
    const rf_scene*               scene,
    const int32_t                 device,
    rf_renderer**                 out)
{
    if (!desc || !scene || !out) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_create: null argument");
    if (desc->max_framebuffer_width <= 0 || desc->max_framebuffer_height <= 0)
        return setError(RF_ERROR_INVALID_ARGUMENT, "maxFramebufferSize must be positive.");
    if (!scene->bvh_nodes || !scene->position_attributes || !scene->vertex_attributes || !scene->base_color_textures ||
        scene->num_base_color_textures == 0)
        return setError(RF_ERROR_INVALID_ARGUMENT, "Scene spans must be non-empty.");
    if (scene->num_position_attributes != scene->num_vertex_attributes)
        return setError(RF_ERROR_INVALID_ARGUMENT, "positionAttributes and vertexAttributes must have the same length.");
    bool          ordered = true;
    std::uint32_t stackEntries = RF_STACK_SIZE;
    rf_status     st = validateBvh(scene->bvh_nodes, scene->num_bvh_nodes, scene->num_position_attributes, ordered, &stackEntries);
    if (st != RF_OK) return st;
    st = validateParams(desc->render_params, desc->max_framebuffer_width, desc->max_framebuffer_height);
    if (st != RF_OK) return st;

    auto r = std::make_unique<rf_renderer>();
    r->stackEntries = stackEntries;
    st = selectDevice(device, r->device, r->numSms);
    if (st != RF_OK) return st;
    r->ordered = ordered;

    // Texture descriptors + concatenated texels (reference_path_tracer.cpp:209-270).
    std::vector<uint4> descs;
    std::uint64_t      totalTexels = 0;
    for (std::uint64_t i = 0; i < scene->num_base_color_textures; ++i)
    {
        const rf_texture& t = scene->base_color_textures[i];
        if (!t.pixels || t.width == 0 || t.height == 0) return setError(RF_ERROR_INVALID_ARGUMENT, "Texture %llu is empty.", (unsigned long long)i);
        descs.push_back(make_uint4(t.width, t.height, static_cast<std::uint32_t>(totalTexels), 0u));
        totalTexels += static_cast<std::uint64_t>(t.width) * t.height;
    }
    const std::uint64_t textureBytes = totalTexels * 4;
    const std::uint64_t maxBinding = 1ull << 30; // REQUIRED_LIMITS.maxStorageBufferBindingSize, pt/gpu_limits.hpp:20-25
    if (textureBytes > maxBinding)
        return setError(
            RF_ERROR_INVALID_ARGUMENT,
            "Texture buffer size (%llu) exceeds maxStorageBufferBindingSize (%llu).",
            (unsigned long long)textureBytes,
            (unsigned long long)maxBinding);

    const std::uint64_t numNodes = scene->num_bvh_nodes, numTris = scene->num_position_attributes;
    {
        // Upload in the reference layouts, then repack on the device.
        DeviceBuffer<rf_bvh_node> rawNodes;
        DeviceBuffer<float>       rawTris;
        RF_CUDA(rawNodes.allocate(numNodes));
        RF_CUDA(rawTris.allocate(numTris * 12));
        RF_CUDA(cudaMemcpy(rawNodes.ptr, scene->bvh_nodes, numNodes * sizeof(rf_bvh_node), cudaMemcpyHostToDevice));
        RF_CUDA(cudaMemcpy(rawTris.ptr, scene->position_attributes, numTris * sizeof(rf_position_attribute), cudaMemcpyHostToDevice));
        // + 64 zeroed records: the tail kernel reads 32-node windows that may run past the last node (straggler.cuh)
        RF_CUDA(r->nodes.allocate(numNodes + 64));
        RF_CUDA(cudaMemset(r->nodes.ptr, 0, (numNodes + 64) * sizeof(PackedNode)));
        RF_CUDA(r->tris.allocate(TRI_STRIDE * numTris));
        k_pack_nodes<<<r->numSms * 4, 256>>>(rawNodes.ptr, numNodes, r->nodes.ptr);
        k_pack_triangles<<<r->numSms * 4, 256>>>(rawTris.ptr, 4, numTris, r->tris.ptr);
        RF_CUDA(cudaGetLastError());
        RF_CUDA(cudaDeviceSynchronize());
    }
    {
        const PairScene ps = buildPairRecords(scene->bvh_nodes, numNodes);
        if (ps.usable)
        {
            RF_CUDA(r->pairRecords.allocate(ps.records.size() + 1)); // (+1: a single-leaf tree has no record)
            if (!ps.records.empty())
                RF_CUDA(cudaMemcpy(r->pairRecords.ptr, ps.records.data(), ps.records.size() * sizeof(PairRecord), cudaMemcpyHostToDevice));
            r->pairsDev.records = r->pairRecords.ptr;
            std::memcpy(r->pairsDev.rootBox, ps.rootBox, sizeof(ps.rootBox));
            r->pairsDev.rootLink = ps.rootLink;
        }
    }
    RF_CUDA(r->vattr.allocate(5 * numTris));
    RF_CUDA(cudaMemcpy(r->vattr.ptr, scene->vertex_attributes, numTris * sizeof(rf_vertex_attributes), cudaMemcpyHostToDevice));
    RF_CUDA(r->texDesc.allocate(descs.size()));
    RF_CUDA(cudaMemcpy(r->texDesc.ptr, descs.data(), descs.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    RF_CUDA(r->texels.allocate(totalTexels));
    for (std::uint64_t i = 0; i < scene->num_base_color_textures; ++i)
    {
        const rf_texture& t = scene->base_color_textures[i];
        RF_CUDA(cudaMemcpy(r->texels.ptr + descs[i].z, t.pixels, static_cast<std::uint64_t>(t.width) * t.height * 4, cudaMemcpyHostToDevice));
    }
    r->numTextures = static_cast<std::uint32_t>(descs.size());
    r->numTexels = totalTexels;

    RF_CUDA(r->blueNoise.allocate(BLUE_NOISE_WIDTH * BLUE_NOISE_HEIGHT));
    RF_CUDA(cudaMemcpy(r->blueNoise.ptr, rf_blue_noise_rg8, BLUE_NOISE_WIDTH * BLUE_NOISE_HEIGHT * 2, cudaMemcpyHostToDevice));
    float srgb[256];
    buildSrgbLut(srgb);
    RF_CUDA(r->srgbLut.allocate(256));
    RF_CUDA(cudaMemcpy(r->srgbLut.ptr, srgb, sizeof(srgb), cudaMemcpyHostToDevice));

    // Frame buffers sized for maxFramebufferSize (reference_path_tracer.cpp:186-190).
    r->maxW = static_cast<std::uint32_t>(desc->max_framebuffer_width);
    r->maxH = static_cast<std::uint32_t>(desc->max_framebuffer_height);
    const std::uint64_t maxPixels = static_cast<std::uint64_t>(r->maxW) * r->maxH;
    const std::uint64_t maxTiles = static_cast<std::uint64_t>((r->maxW + TILE - 1) / TILE) * ((r->maxH + TILE - 1) / TILE);
    (void)maxTiles;
    RF_CUDA(r->image.allocate(maxPixels));
    RF_CUDA(r->radiance.allocate(maxPixels));
    RF_CUDA(r->display.allocate(maxPixels));
    RF_CUDA(r->stats.allocate(STAT_COUNT));
    RF_CUDA(cudaMemset(r->image.ptr, 0, maxPixels * sizeof(float4)));
    RF_CUDA(cudaMemset(r->stats.ptr, 0, STAT_COUNT * sizeof(unsigned long long)));
    RF_CUDA(cudaEventCreateWithFlags(&r->forkEvent, cudaEventDisableTiming));
    for (int i = 0; i < rf_renderer::MAX_SUBFRAMES; ++i)
    {
        if (i > 0) RF_CUDA(cudaStreamCreateWithFlags(&r->sub[i].stream, cudaStreamNonBlocking));
        RF_CUDA(cudaEventCreateWithFlags(&r->sub[i].done, cudaEventDisableTiming));
    }

    st = r->applyParams(desc->render_params);
    if (st != RF_OK) return st;
    *out = r.release();
    return RF_OK;
}

extern "C" void rf_renderer_destroy(rf_renderer* r) { delete r; }

extern "C" rf_status rf_renderer_set_render_parameters(rf_renderer* r, const rf_render_parameters* params)
{
    if (!r || !params) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_render_parameters: null argument");
    if (sameParams(r->params, *params)) return RF_OK; // reference_path_tracer.cpp:558
    const rf_status st = validateParams(*params, r->maxW, r->maxH);
    if (st != RF_OK) return st;
    RF_CUDA(cudaSetDevice(r->device));
    return r->applyParams(*params);
}

extern "C" rf_status rf_renderer_set_tile_partition(rf_renderer* r, std::uint32_t rank, std::uint32_t world)
{
    if (!r || world == 0 || rank >= world) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_tile_partition: need rank < world, world >= 1");
    r->rank = rank, r->world = world;
    r->tilesDirty = true;
    r->accumulated = 0;
    return RF_OK;
}

extern "C" rf_status rf_renderer_render(rf_renderer* r)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_render: null renderer");
    RF_CUDA(cudaSetDevice(r->device));
    const std::uint32_t spp = r->params.sampling_params.num_samples_per_pixel;
    const std::uint32_t frameCount = r->frameCount++; // mFrameCount++ (reference_path_tracer.cpp:580)
    if (r->accumulated >= spp) return RF_OK;          // fsMain:51 — nothing is traced once converged

    rf_status st = r->updateTiles();
    if (st != RF_OK) return st;
    r->drainTimings(false);

    cudaStream_t s = r->stream;
    FrameParams  fp{};
    fp.width = r->params.framebuffer_width;
    fp.height = r->params.framebuffer_height;
    fp.frameCount = frameCount;
    fp.sampleIndex = frameCount % spp;
    fp.numBounces = r->params.sampling_params.num_bounces;
    fp.tilesX = r->tilesX;
    fp.numTextures = r->numTextures;
    fp.numTexels = r->numTexels;
    fp.camera = r->params.camera;
    fp.sky = r->skyState;
    const SolarConstants sc = solarConstants();
    fp.solarCosThetaMax = sc.cosThetaMax;
    fp.solarInvPdf = sc.invPdf;

    SceneDevice scene{r->nodes.ptr, r->tris.ptr, r->vattr.ptr, r->texDesc.ptr, r->texels.ptr, r->blueNoise.ptr, r->lut.ptr, r->srgbLut.ptr,
                      r->ordered, r->tuning, r->pairsDev};

    rf_renderer::Timed t{};
    if (!r->eventPool.empty())
    {
        t = std::move(r->eventPool.back());
        r->eventPool.pop_back();
    }
    else
    {
        RF_CUDA(cudaEventCreate(&t.begin));
        RF_CUDA(cudaEventCreate(&t.end));
    }
    t.stagesUsed = 0;
    t.bounces = fp.numBounces;
    // Per-stage events are only meaningful when the stages do not overlap: one sub-frame on one stream.
    const bool staged = r->stageTiming && r->numSubFrames == 1; // (also selects the staged pipeline: one launch per stage)
    const auto stageMark = [&]() -> cudaError_t {
        if (!staged) return cudaSuccess;
        if (t.stagesUsed >= t.stages.size())
        {
            cudaEvent_t e;
            const cudaError_t err = cudaEventCreate(&e);
            if (err != cudaSuccess) return err;
            t.stages.push_back(e);
        }
        return cudaEventRecord(t.stages[t.stagesUsed++], s);
    };

    RF_CUDA(cudaEventRecord(t.begin, s));
    const std::uint64_t numPixels = static_cast<std::uint64_t>(fp.width) * fp.height;
    const bool restart = r->accumulated == 0;
    if (restart)
    {
        // fsMain:45-47: imageBuffer[idx] = vec3(0f) on the first sample; k_accumulate does that for the owned pixels, this
        // clears the others (what a sum-reduce over ranks needs).  The local image is never written by another rank (the
        // peer-memory exchange has its own double-buffered target), so the clear cannot race with anything.
        RF_CUDA(cudaMemsetAsync(r->image.ptr, 0, numPixels * sizeof(float4), s));
    }
    float4* const exchangeTarget = r->exchangeTarget();
    RF_CUDA(cudaEventRecord(r->forkEvent, s));

    const int gridLight = r->gridFor(8);
    const int gridTrace = r->usePairs() && !r->useMega() ? r->gridFor(r->effectiveBlocksPerSm()) : r->gridFor(r->effectiveBlocksPerSm() * (256 / r->traceBlock));
    RF_CUDA(stageMark());
    for (int i = 0; i < r->numSubFrames; ++i)
    {
        rf_renderer::SubFrame& sf = r->sub[i];
        if (sf.numOwnedTiles == 0) continue;
        cudaStream_t ss = i == 0 ? s : sf.stream;
        if (i > 0) RF_CUDA(cudaStreamWaitEvent(ss, r->forkEvent, 0));
        FrameParams sfp = fp;
        sfp.numOwnedTiles = sf.numOwnedTiles;
        std::uint32_t* ctr = sf.counters.ptr;
        RF_CUDA(cudaMemsetAsync(ctr, 0, counterSlots(fp.numBounces) * sizeof(std::uint32_t), ss));
        std::uint32_t* const cursors = ctr + fp.numBounces + 1u;
        if (r->useMega() && !staged)
        {
            // One launch for the tile set's whole frame (mega.cuh): all SM slots divided among the tile sets, and as many path
            // slots per block as the block's share of the pixels needs (every path in flight from the start: the paths of a
            // block then advance together and end together), at most loopMaxSlots().
            const int           megaGrid = r->gridFor(r->traceBlocksPerSm > 0 ? r->traceBlocksPerSm : std::max(1, 4 / r->numSubFrames)) * 256 / r->megaBlock;
            const std::uint64_t pathsPerBlock = (static_cast<std::uint64_t>(sf.numOwnedTiles) * TILE_PIXELS + megaGrid - 1) / megaGrid;
            const std::uint32_t maxSlots = loopMaxSlots(r->megaBlock);
            // Path slots per block.  Paths that cannot start at once start as others end, in "waves"; a last wave that is only
            // partly filled leaves lanes idle for a whole path's latency (measured at 960x540: 1760 slots for 1750 paths per block
            // 4.37 ms, 1536 slots 5.05 ms), so the share is split into the fewest waves that fit, all of the same size.
            const std::uint64_t waves = (pathsPerBlock + maxSlots - 1) / maxSlots;
            const std::uint32_t slots = r->megaSlots != 0u ? std::min(r->megaSlots, maxSlots)
                                                           : static_cast<std::uint32_t>(std::max<std::uint64_t>(64u, ((pathsPerBlock + waves - 1) / waves + 31u) & ~31ull));
            const std::uint64_t recordSlots = static_cast<std::uint64_t>(megaGrid) * slots;
            if (recordSlots > sf.pathRecordSlots)
            {
                RF_CUDA(cudaStreamSynchronize(ss));
                RF_CUDA(sf.pathRecords.allocate(recordSlots * LOOP_RECORD_VEC));
                sf.pathRecordSlots = recordSlots;
            }
            launchMega(r->variant, r->megaBlock, r->traceStackEntries(), megaGrid, ss, sfp, scene, sf.ownedTiles.ptr, sf.pathRecords.ptr, slots, r->radiance.ptr, &ctr[0],
                       reinterpret_cast<std::uint32_t*>(r->stats.ptr + STAT_FAILED), r->stats.ptr);
            k_accumulate<<<gridLight, BLOCK_THREADS, 0, ss>>>(sfp, sf.ownedTiles.ptr, r->radiance.ptr, r->image.ptr, exchangeTarget, restart);
            r->kernelLaunches += 2;
            if (i > 0)
            {
                RF_CUDA(cudaEventRecord(sf.done, ss));
                RF_CUDA(cudaStreamWaitEvent(s, sf.done, 0));
            }
            continue;
        }

        // Straggler hand-over of the k-th traversal launch (off in the staged timing mode: one launch per stage).
        const std::uint32_t evictMax = staged ? 0u : r->effectiveEvictMax();
        std::uint32_t* const stragglerCounts = ctr + 2u * fp.numBounces + 2u;
        std::uint32_t* const stragglerCursors = ctr + 3u * fp.numBounces + 3u;
        const std::uint32_t evictDelay = r->evictDelay;
        const auto           stragglersOf = [&](std::uint32_t k) { return StragglerBuffer{sf.stragglers.ptr, &stragglerCounts[k], r->stragglerCapacity(), evictMax, evictDelay}; };
        const auto           finishStragglers = [&](std::uint32_t k, const PathQueue& closestQueue, const PathQueue& shadowQueue) {
            if (evictMax == 0u) return;
            // persistent warps, one ray at a time each; blocks beyond the number of records exit at once
            const int tailGrid = r->numSms * STRAGGLER_WARPS_PER_SM / STRAGGLER_WARPS_PER_BLOCK;
            if (r->stragglerWindowMode == STRAGGLER_BULK)
                k_trace_stragglers<STRAGGLER_BULK><<<tailGrid, STRAGGLER_BLOCK_THREADS, 0, ss>>>(
                    sfp, scene, closestQueue, sf.hits.ptr, shadowQueue, r->radiance.ptr, &stragglerCursors[k], stragglersOf(k), r->stats.ptr);
            else
                k_trace_stragglers<STRAGGLER_DIRECT><<<tailGrid, STRAGGLER_BLOCK_THREADS, 0, ss>>>(
                    sfp, scene, closestQueue, sf.hits.ptr, shadowQueue, r->radiance.ptr, &stragglerCursors[k], stragglersOf(k), r->stats.ptr);
        };

        k_raygen<<<gridLight, BLOCK_THREADS, 0, ss>>>(sfp, scene, sf.ownedTiles.ptr, sf.queues[0], &ctr[0], r->radiance.ptr, r->stats.ptr);
        RF_CUDA(stageMark());
        // closest-hit rays of bounce 1
        const bool pairsKernel = r->usePairs();
        if (pairsKernel)
            launchTracePairs(r->pairVariant, gridTrace, ss, sfp, scene, sf.queues[0], &ctr[0], sf.hits.ptr, sf.queues[0], nullptr, r->radiance.ptr, &cursors[0], r->stats.ptr);
        else
            launchTrace(r->variant, r->traceBlock, r->traceStackEntries(), gridTrace, ss, sfp, scene, sf.queues[0], &ctr[0], sf.hits.ptr, sf.queues[0], nullptr, r->radiance.ptr, &cursors[0],
                        stragglersOf(0), r->stats.ptr);
        finishStragglers(0, sf.queues[0], sf.queues[0]);
        for (std::uint32_t bounce = 1; bounce <= fp.numBounces; ++bounce)
        {
            const int in = (bounce - 1) & 1, outQ = bounce & 1;
            RF_CUDA(stageMark());
            k_shade<<<gridLight, BLOCK_THREADS, 0, ss>>>(sfp, scene, sf.queues[in], &ctr[bounce - 1], sf.hits.ptr, sf.queues[outQ], &ctr[bounce], r->radiance.ptr);
            RF_CUDA(stageMark());
            // shadow rays of this bounce + closest-hit rays of the next one (none after the last bounce)
            const bool last = bounce == fp.numBounces;
            if (pairsKernel)
                launchTracePairs(r->pairVariant, gridTrace, ss, sfp, scene, sf.queues[outQ], last ? nullptr : &ctr[bounce], sf.hits.ptr, sf.queues[outQ], &ctr[bounce],
                                 r->radiance.ptr, &cursors[bounce], r->stats.ptr);
            else
                launchTrace(r->variant, r->traceBlock, r->traceStackEntries(), gridTrace, ss, sfp, scene, sf.queues[outQ], last ? nullptr : &ctr[bounce], sf.hits.ptr, sf.queues[outQ], &ctr[bounce],
                            r->radiance.ptr, &cursors[bounce], stragglersOf(bounce), r->stats.ptr);
            finishStragglers(bounce, sf.queues[outQ], sf.queues[outQ]);
        }
        RF_CUDA(stageMark());
        k_accumulate<<<gridLight, BLOCK_THREADS, 0, ss>>>(sfp, sf.ownedTiles.ptr, r->radiance.ptr, r->image.ptr, exchangeTarget, restart);
        // raygen + (numBounces + 1) traversal launches (+ their straggler follow-ups) + numBounces shades + accumulate
        r->kernelLaunches += 2ull + (fp.numBounces + 1ull) * (evictMax != 0u ? 2ull : 1ull) + fp.numBounces;
        RF_CUDA(stageMark());
        if (i > 0)
        {
            RF_CUDA(cudaEventRecord(sf.done, ss));
            RF_CUDA(cudaStreamWaitEvent(s, sf.done, 0));
        }
    }
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaEventRecord(t.end, s));
    r->pending.push_back(std::move(t));
    r->frames++;
    r->accumulated = std::min(r->accumulated + 1, spp); // reference_path_tracer.cpp:590-591
    if (r->peerImage) ++r->exchangeEpoch;

    return RF_OK;
}

// DeferredRenderer::render's lighting pass + resolve pass (pt/deferred_renderer.cpp:340-375) on a caller-supplied
// G-buffer; see deferred.cuh.
extern "C" rf_status rf_renderer_render_deferred_lighting(
    rf_renderer*                       r,
    const rf_deferred_lighting_params* p,
    const float*                       gbufferAlbedo,
    const float*                       gbufferNormal,
    const float*                       gbufferDepth)
{
    if (!r || !p || !gbufferAlbedo || !gbufferNormal || !gbufferDepth)
        return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_render_deferred_lighting: null argument");
    if (p->framebuffer_width == 0 || p->framebuffer_height == 0 || p->framebuffer_width > r->maxW || p->framebuffer_height > r->maxH)
        return setError(RF_ERROR_INVALID_ARGUMENT, "Framebuffer size %ux%u outside (0, %ux%u].", p->framebuffer_width, p->framebuffer_height, r->maxW, r->maxH);
    RF_CUDA(cudaSetDevice(r->device));
    rf_sky_state    sky{};
    const rf_status st = rf_sky_state_new(&p->sky, &sky);
    if (st != RF_OK) return st;

    rf_renderer::Deferred& d = r->deferred;
    const std::uint32_t    w = p->framebuffer_width, h = p->framebuffer_height;
    const std::uint64_t    numPixels = static_cast<std::uint64_t>(w) * h;
    cudaStream_t           s = r->stream;
    if (numPixels > d.capacity)
    {
        RF_CUDA(cudaStreamSynchronize(s));
        RF_CUDA(d.albedo.allocate(numPixels));
        RF_CUDA(d.normal.allocate(numPixels));
        RF_CUDA(d.depth.allocate(numPixels));
        RF_CUDA(d.accumulation.allocate(numPixels));
        RF_CUDA(d.queueMem.allocate(numPixels * 8));
        RF_CUDA(d.hits.allocate(numPixels));
        if (!d.counters.ptr) RF_CUDA(d.counters.allocate(counterSlots(1)));
        if (!d.lutRow.ptr) RF_CUDA(d.lutRow.allocate(1));
        d.capacity = numPixels;
    }
    if (w != d.width || h != d.height) RF_CUDA(cudaMemsetAsync(d.accumulation.ptr, 0, numPixels * sizeof(float4), s));
    d.width = w, d.height = h, d.exposure = p->exposure;
    PathQueue queues[2];
    for (int q = 0; q < 2; ++q)
    {
        float4* base = d.queueMem.ptr + static_cast<std::uint64_t>(q) * 4 * numPixels;
        queues[q] = PathQueue{base, base + numPixels, base + 2 * numPixels, base + 3 * numPixels};
    }

    // host -> device: the G-buffer of this frame and the sampling table of animatedBlueNoise(coord, frameCount, 1 << 20)
    RF_CUDA(cudaMemcpyAsync(d.albedo.ptr, gbufferAlbedo, numPixels * sizeof(float4), cudaMemcpyHostToDevice, s));
    RF_CUDA(cudaMemcpyAsync(d.normal.ptr, gbufferNormal, numPixels * sizeof(float4), cudaMemcpyHostToDevice, s));
    RF_CUDA(cudaMemcpyAsync(d.depth.ptr, gbufferDepth, numPixels * sizeof(float), cudaMemcpyHostToDevice, s));
    SampleLutRow row;
    buildSampleLutRow(p->frame_count % (1u << 20), row);
    RF_CUDA(cudaMemcpyAsync(d.lutRow.ptr, &row, sizeof(row), cudaMemcpyHostToDevice, s));
    RF_CUDA(cudaStreamSynchronize(s)); // `row` and the caller's buffers may go away

    FrameParams fp{};
    fp.width = w, fp.height = h;
    fp.frameCount = p->frame_count;
    fp.sampleIndex = 0; // the one row of d.lutRow
    fp.numBounces = 1;  // NUM_BOUNCES = 2 counts the G-buffer surface
    fp.numTextures = r->numTextures;
    fp.numTexels = r->numTexels;
    fp.sky = sky;
    const SolarConstants sc = solarConstants();
    fp.solarCosThetaMax = sc.cosThetaMax;
    fp.solarInvPdf = sc.invPdf;
    fp.deferred = 1u;
    SceneDevice scene{r->nodes.ptr, r->tris.ptr, r->vattr.ptr, r->texDesc.ptr, r->texels.ptr, r->blueNoise.ptr, d.lutRow.ptr, r->srgbLut.ptr,
                      r->ordered, r->tuning, r->pairsDev};
    DeferredUniforms un{};
    std::memcpy(un.inverseViewReverseZProjection, p->inverse_view_reverse_z_projection, sizeof(un.inverseViewReverseZProjection));
    std::memcpy(un.cameraEye, p->camera_eye, sizeof(un.cameraEye));
    un.frameCount = p->frame_count;

    // device time of the passes (without the G-buffer upload), reported like the path tracer's render-pass duration
    r->drainTimings(false);
    rf_renderer::Timed t{};
    if (!r->eventPool.empty())
    {
        t = std::move(r->eventPool.back());
        r->eventPool.pop_back();
    }
    else
    {
        RF_CUDA(cudaEventCreate(&t.begin));
        RF_CUDA(cudaEventCreate(&t.end));
    }
    t.stagesUsed = 0;
    t.bounces = 1;
    RF_CUDA(cudaEventRecord(t.begin, s));

    std::uint32_t* ctr = d.counters.ptr;
    std::uint32_t* const cursors = ctr + 2;
    const StragglerBuffer noHandOver{nullptr, nullptr, 0u, 0u, 0u};
    const int gridLight = r->gridFor(8);
    const int gridTrace = r->usePairs() && !r->useMega() ? r->gridFor(r->effectiveBlocksPerSm()) : r->gridFor(r->effectiveBlocksPerSm() * (256 / r->traceBlock));
    RF_CUDA(cudaMemsetAsync(ctr, 0, counterSlots(1) * sizeof(std::uint32_t), s));
    k_deferred_primary<<<gridLight, BLOCK_THREADS, 0, s>>>(fp, scene, un, d.albedo.ptr, d.normal.ptr, d.depth.ptr, queues[0], &ctr[0], r->radiance.ptr, r->stats.ptr);
    // shadow rays of the G-buffer surfaces + the bounce rays
    if (r->usePairs())
        launchTracePairs(r->pairVariant, gridTrace, s, fp, scene, queues[0], &ctr[0], d.hits.ptr, queues[0], &ctr[0], r->radiance.ptr, &cursors[0], r->stats.ptr);
    else
        launchTrace(r->variant, r->traceBlock, r->traceStackEntries(), gridTrace, s, fp, scene, queues[0], &ctr[0], d.hits.ptr, queues[0], &ctr[0], r->radiance.ptr, &cursors[0], noHandOver, r->stats.ptr);
    k_shade<<<gridLight, BLOCK_THREADS, 0, s>>>(fp, scene, queues[0], &ctr[0], d.hits.ptr, queues[1], &ctr[1], r->radiance.ptr);
    // shadow rays of the bounce hits
    if (r->usePairs())
        launchTracePairs(r->pairVariant, gridTrace, s, fp, scene, queues[1], nullptr, d.hits.ptr, queues[1], &ctr[1], r->radiance.ptr, &cursors[1], r->stats.ptr);
    else
        launchTrace(r->variant, r->traceBlock, r->traceStackEntries(), gridTrace, s, fp, scene, queues[1], nullptr, d.hits.ptr, queues[1], &ctr[1], r->radiance.ptr, &cursors[1], noHandOver, r->stats.ptr);
    k_deferred_resolve<<<gridLight, BLOCK_THREADS, 0, s>>>(static_cast<std::uint32_t>(numPixels), p->frame_count, r->radiance.ptr, d.accumulation.ptr);
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaEventRecord(t.end, s));
    r->pending.push_back(std::move(t));
    r->kernelLaunches += 5;
    r->frames++;
    return RF_OK;
}

// sampleBuffer / accumulationBuffer (array<array<f32, 3>>, deferred_renderer_lighting_pass.wgsl:87, resolve pass :31-32) and
// the resolve pass's colour attachment (aces(exposure * colour)^(1/2.2), BGRA8) of the last deferred frame.  Each may be NULL.
extern "C" rf_status rf_renderer_read_deferred(rf_renderer* r, float* sampleRgb, float* accumulationRgb, std::uint32_t* displayBgra8)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_read_deferred: null renderer");
    rf_renderer::Deferred& d = r->deferred;
    if (d.width == 0) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_read_deferred: no deferred frame rendered yet");
    RF_CUDA(cudaSetDevice(r->device));
    const std::uint64_t numPixels = static_cast<std::uint64_t>(d.width) * d.height;
    std::vector<float4> tmp(numPixels);
    const auto          fetch = [&](const float4* src, float* dst) -> cudaError_t {
        const cudaError_t err = cudaMemcpyAsync(tmp.data(), src, numPixels * sizeof(float4), cudaMemcpyDeviceToHost, r->stream);
        if (err != cudaSuccess) return err;
        const cudaError_t err2 = cudaStreamSynchronize(r->stream);
        for (std::uint64_t i = 0; i < numPixels && err2 == cudaSuccess; ++i) dst[3 * i] = tmp[i].x, dst[3 * i + 1] = tmp[i].y, dst[3 * i + 2] = tmp[i].z;
        return err2;
    };
    if (sampleRgb) RF_CUDA(fetch(r->radiance.ptr, sampleRgb));
    if (accumulationRgb) RF_CUDA(fetch(d.accumulation.ptr, accumulationRgb));
    if (displayBgra8)
    {
        k_display<<<r->gridFor(8), BLOCK_THREADS, 0, r->stream>>>(static_cast<std::uint32_t>(numPixels), d.accumulation.ptr, 1.0f, d.exposure, r->display.ptr);
        RF_CUDA(cudaGetLastError());
        RF_CUDA(cudaMemcpyAsync(displayBgra8, r->display.ptr, numPixels * 4, cudaMemcpyDeviceToHost, r->stream));
        RF_CUDA(cudaStreamSynchronize(r->stream));
    }
    return RF_OK;
}

extern "C" float rf_renderer_average_renderpass_duration_ms(rf_renderer* r)
{
    if (!r) return 0.0f;
    cudaSetDevice(r->device);
    r->drainTimings(false);
    if (r->durationsMs.empty()) return 0.0f; // reference_path_tracer.cpp:708-711
    double sum = 0.0;
    for (float ms : r->durationsMs) sum += ms;
    return static_cast<float>(sum / static_cast<double>(r->durationsMs.size()));
}

extern "C" float rf_renderer_render_progress_percentage(const rf_renderer* r)
{
    if (!r) return 0.0f;
    return 100.0f * static_cast<float>(r->accumulated) / static_cast<float>(r->params.sampling_params.num_samples_per_pixel);
}

extern "C" rf_status rf_renderer_read_hdr(rf_renderer* r, float* dst, std::uint64_t numFloats, std::uint32_t* accumulated)
{
    if (!r || !dst) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_read_hdr: null argument");
    const std::uint64_t need = static_cast<std::uint64_t>(r->params.framebuffer_width) * r->params.framebuffer_height * 4;
    if (numFloats < need) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_read_hdr: need %llu floats, got %llu", (unsigned long long)need, (unsigned long long)numFloats);
    RF_CUDA(cudaSetDevice(r->device));
    RF_CUDA(cudaMemcpyAsync(dst, r->presentedImage(), need * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
    RF_CUDA(cudaStreamSynchronize(r->stream));
    if (accumulated) *accumulated = r->accumulated;
    return RF_OK;
}

extern "C" rf_status rf_renderer_read_display(rf_renderer* r, std::uint32_t* dst, std::uint64_t numPixels)
{
    if (!r || !dst) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_read_display: null argument");
    const std::uint64_t need = static_cast<std::uint64_t>(r->params.framebuffer_width) * r->params.framebuffer_height;
    if (numPixels < need) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_read_display: need %llu pixels", (unsigned long long)need);
    RF_CUDA(cudaSetDevice(r->device));
    const float acc = static_cast<float>(std::max(r->accumulated, 1u));
    k_display<<<r->gridFor(8), BLOCK_THREADS, 0, r->stream>>>(static_cast<std::uint32_t>(need), r->presentedImage(), acc, r->params.exposure, r->display.ptr);
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaMemcpyAsync(dst, r->display.ptr, need * 4, cudaMemcpyDeviceToHost, r->stream));
    RF_CUDA(cudaStreamSynchronize(r->stream));
    return RF_OK;
}

extern "C" void* rf_renderer_hdr_device_ptr(rf_renderer* r) { return r ? r->image.ptr : nullptr; }

extern "C" rf_status rf_renderer_set_stream(rf_renderer* r, void* cudaStream)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_stream: null renderer");
    RF_CUDA(cudaSetDevice(r->device));
    RF_CUDA(cudaStreamSynchronize(r->stream));
    r->stream = static_cast<cudaStream_t>(cudaStream);
    return RF_OK;
}

extern "C" rf_status rf_renderer_synchronize(rf_renderer* r)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_synchronize: null renderer");
    RF_CUDA(cudaSetDevice(r->device));
    RF_CUDA(cudaStreamSynchronize(r->stream));
    r->drainTimings(true);
    return RF_OK;
}

extern "C" rf_status rf_renderer_set_frame_count(rf_renderer* r, std::uint32_t frameCount)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_frame_count: null renderer");
    r->frameCount = frameCount;
    return RF_OK;
}
extern "C" std::uint32_t rf_renderer_frame_count(const rf_renderer* r) { return r ? r->frameCount : 0; }
extern "C" std::uint32_t rf_renderer_accumulated_sample_count(const rf_renderer* r) { return r ? r->accumulated : 0; }

extern "C" rf_status rf_renderer_get_stats(rf_renderer* r, rf_frame_stats* out)
{
    if (!r || !out) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_get_stats: null argument");
    RF_CUDA(cudaSetDevice(r->device));
    RF_CUDA(cudaStreamSynchronize(r->stream));
    r->drainTimings(true);
    unsigned long long s[STAT_COUNT];
    RF_CUDA(cudaMemcpy(s, r->stats.ptr, sizeof(s), cudaMemcpyDeviceToHost));
    std::memset(out, 0, sizeof(*out));
    out->frames = r->frames;
    out->paths = s[STAT_PATHS];
    out->closest_rays = s[STAT_CLOSEST_RAYS];
    out->shadow_rays = s[STAT_SHADOW_RAYS];
    out->closest_nodes_visited = s[STAT_CLOSEST_NODES];
    out->closest_triangles_tested = s[STAT_CLOSEST_TRIS];
    out->shadow_nodes_visited = s[STAT_SHADOW_NODES];
    out->shadow_triangles_tested = s[STAT_SHADOW_TRIS];
    out->device_ms_total = r->totalMs;
    out->device_ms_trace = r->msTrace;
    out->device_ms_shade = r->msShade;
    out->device_ms_other = r->msOther;
    out->kernel_launches = r->kernelLaunches;
    out->sub_frames = static_cast<std::uint32_t>(r->numSubFrames);
    out->evict_max = r->useMega() ? 0u : r->effectiveEvictMax();
    out->node_records_loaded = s[STAT_RECORDS];
    out->trace_kernel = r->usePairs() && !r->useMega() ? 2u : 1u;
    out->persistent_kernel = r->useMega() ? 1u : 0u;
    if (s[STAT_FAILED] != 0ull) return setError(RF_ERROR_CUDA, "The persistent kernel left a frame on its watchdog (a lost path); the image is incomplete.");
    return RF_OK;
}

extern "C" rf_status rf_renderer_reset_stats(rf_renderer* r)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_reset_stats: null renderer");
    RF_CUDA(cudaSetDevice(r->device));
    RF_CUDA(cudaStreamSynchronize(r->stream));
    r->drainTimings(true);
    RF_CUDA(cudaMemset(r->stats.ptr, 0, STAT_COUNT * sizeof(unsigned long long)));
    r->frames = 0;
    r->kernelLaunches = 0;
    r->totalMs = r->msTrace = r->msShade = r->msOther = 0.0;
    return RF_OK;
}

extern "C" rf_status rf_renderer_set_stage_timing(rf_renderer* r, int32_t enabled)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_stage_timing: null renderer");
    r->stageTiming = enabled != 0;
    return RF_OK;
}

extern "C" rf_status rf_renderer_set_tuning(rf_renderer* r, std::uint32_t triMin, std::uint32_t refillMin, std::uint32_t blocksPerSm)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_tuning: null renderer");
    if (triMin > 32u || refillMin > 32u || blocksPerSm > 8u) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_tuning: value out of range");
    if (triMin) r->tuning.triMin = triMin;
    if (refillMin) r->tuning.refillMin = refillMin;
    if (blocksPerSm) r->traceBlocksPerSm = static_cast<int>(blocksPerSm);
    return RF_OK;
}

#ifdef RF_TRACE_TIMELINE
// Debug build only: arm the per-warp timeline of the traversal launches / read it back (tools/trace_timeline.py).
extern "C" int rf_debug_timeline_arm(std::uint32_t capacity)
{
    static TimelineRecord* buf = nullptr;
    if (buf) cudaFree(buf);
    if (cudaMalloc(&buf, sizeof(TimelineRecord) * capacity) != cudaSuccess) return 1;
    const std::uint32_t zero = 0;
    cudaMemcpyToSymbol(g_timeline, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_timelineCount, &zero, sizeof(zero));
    cudaMemcpyToSymbol(g_timelineCap, &capacity, sizeof(capacity));
    static const unsigned long long zeros[OCC_BUCKETS] = {};
    cudaMemcpyToSymbol(g_occBusy, zeros, sizeof(zeros));
    cudaMemcpyToSymbol(g_occRounds, zeros, sizeof(zeros));
    return cudaDeviceSynchronize() != cudaSuccess;
}
// busy[256] then rounds[256]: lane occupancy of the traversal warps per 16.4 us bucket of %globaltimer (traversal.cuh)
extern "C" void rf_debug_occupancy_read(unsigned long long* dst)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(dst, g_occBusy, sizeof(unsigned long long) * OCC_BUCKETS);
    cudaMemcpyFromSymbol(dst + OCC_BUCKETS, g_occRounds, sizeof(unsigned long long) * OCC_BUCKETS);
}
extern "C" std::uint32_t rf_debug_timeline_read(void* dst, std::uint32_t capacity)
{
    cudaDeviceSynchronize();
    std::uint32_t   n = 0;
    TimelineRecord* buf = nullptr;
    cudaMemcpyFromSymbol(&n, g_timelineCount, sizeof(n));
    cudaMemcpyFromSymbol(&buf, g_timeline, sizeof(buf));
    n = std::min(n, capacity);
    if (buf && n) cudaMemcpy(dst, buf, sizeof(TimelineRecord) * n, cudaMemcpyDeviceToHost);
    return n;
}
#endif

// ---- multi-GPU exchange over NVLink peer memory -----------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "rf_renderer_hdr_ipc_handle hands out 64 bytes");

extern "C" rf_status rf_renderer_hdr_ipc_handle(rf_renderer* r, void* outHandle64)
{
    if (!r || !outHandle64) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_hdr_ipc_handle: null argument");
    RF_CUDA(cudaSetDevice(r->device));
    RF_CUDA(cudaStreamSynchronize(r->stream));
    if (r->peerImage && !r->exportedRoot) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_hdr_ipc_handle: this renderer is attached to another root");
    if (!r->exchange.ptr)
    {
        RF_CUDA(r->exchange.allocate(2 * r->exchangeStride()));
        RF_CUDA(cudaMemset(r->exchange.ptr, 0, 2 * r->exchangeStride() * sizeof(float4)));
    }
    cudaIpcMemHandle_t h;
    RF_CUDA(cudaIpcGetMemHandle(&h, r->exchange.ptr));
    std::memcpy(outHandle64, &h, sizeof(h));
    r->exportedRoot = true;
    r->peerImage = r->exchange.ptr; // the root stores its own pixels there too
    r->exchangeEpoch = 0;
    return RF_OK;
}

extern "C" rf_status rf_renderer_set_hdr_peer(rf_renderer* r, const void* handle64)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_hdr_peer: null renderer");
    RF_CUDA(cudaSetDevice(r->device));
    RF_CUDA(cudaStreamSynchronize(r->stream));
    for (auto& sf : r->sub)
        if (sf.stream) RF_CUDA(cudaStreamSynchronize(sf.stream));
    if (r->peerImage && !r->exportedRoot) RF_CUDA(cudaIpcCloseMemHandle(r->peerImage));
    r->peerImage = nullptr;
    r->exchangeEpoch = 0;
    if (!handle64)
    {
        r->exportedRoot = false; // the root keeps its buffer until the renderer is destroyed (peers may still unmap it)
        return RF_OK;
    }
    if (r->exportedRoot) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_hdr_peer: this renderer is the root of an exchange");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    void* mapped = nullptr;
    RF_CUDA(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
    r->peerImage = static_cast<float4*>(mapped);
    return RF_OK;
}

extern "C" void* rf_renderer_exchange_device_ptr(rf_renderer* r)
{
    return r && r->exportedRoot ? const_cast<float4*>(r->presentedImage()) : nullptr;
}

extern "C" rf_status rf_renderer_set_tail_policy(rf_renderer* r, std::int32_t evictMax)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_tail_policy: null renderer");
    if (evictMax > 32) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_tail_policy: value out of range");
    r->evictMax = evictMax < 0 ? -1 : evictMax;
    return RF_OK;
}

// Named scheduling / debugging knobs that used to be environment variables; results never depend on them.
extern "C" rf_status rf_renderer_set_option(rf_renderer* r, const char* name, std::int64_t value)
{
    if (!r || !name) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_option: null argument");
    const std::string key(name);
    if (value < 0 || value > (1ll << 30)) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_option: value of '%s' out of range", name);
    if (key == "shade_wait") r->tuning.shadeWait = static_cast<std::uint32_t>(value);
    else if (key == "tail_paths") r->tuning.tailPaths = static_cast<std::uint32_t>(value);
    else if (key == "walk_in_place" && value <= 1) r->tuning.walkInPlace = static_cast<std::uint32_t>(value);
    else if (key == "priority_mode" && value <= 2) r->tuning.priorityMode = static_cast<std::uint32_t>(value);
    else if (key == "evict_delay") r->evictDelay = static_cast<std::uint32_t>(value);
    else if (key == "trace_stack") r->forcedStackEntries = static_cast<std::uint32_t>(std::min<std::int64_t>(value, RF_STACK_SIZE));
    else if (key == "trace_kernel" && value <= 2) r->traceKernel = static_cast<int>(value);
    else if (key == "pair_variant" && value <= 7) r->pairVariant = static_cast<int>(value);
    else if (key == "tail_window_mode" && value <= 1) r->stragglerWindowMode = static_cast<int>(value);
    else if (key == "stage_debug") r->stageDebug = value != 0;
    else if (key == "mega_slots" && (value == 0 || (value >= 64 && value <= 2048 && value % 32 == 0))) r->megaSlots = static_cast<std::uint32_t>(value);
    else if (key == "mega_block" && (value == 256 || value == 512)) r->megaBlock = static_cast<int>(value);
    else return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_option: unknown option '%s'", name);
    return RF_OK;
}

extern "C" rf_status rf_renderer_set_pipeline(rf_renderer* r, std::int32_t subFrames, std::int32_t persistentKernel, std::int32_t variant, std::int32_t blockThreads)
{
    if (!r) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_pipeline: null renderer");
    if (subFrames > rf_renderer::MAX_SUBFRAMES || persistentKernel > 2 || variant > 15 || (blockThreads != 0 && blockThreads != 64 && blockThreads != 128 && blockThreads != 256))
        return setError(RF_ERROR_INVALID_ARGUMENT, "rf_renderer_set_pipeline: value out of range");
    if (subFrames != 0 && (subFrames < 0 ? 0 : subFrames) != r->requestedSubFrames) r->requestedSubFrames = subFrames < 0 ? 0 : subFrames, r->tilesDirty = true;
    if (persistentKernel >= 0 && persistentKernel != r->megaMode) r->megaMode = persistentKernel, r->tilesDirty = true;
    if (variant >= 0) r->variant = variant;
    if (blockThreads > 0) r->traceBlock = blockThreads;
    return RF_OK;
}

// =================================================================================================
struct rf_traversal_scene
{
    int                  device = 0;
    int                  numSms = 0;
    DeviceBuffer<PackedNode>    nodes;
    DeviceBuffer<float4>        tris;
    DeviceBuffer<std::uint32_t> cursor;
    bool                        ordered = true;
    TraceTuning                 tuning = defaultTuning();
    std::uint64_t               numNodes = 0, numTris = 0;
    DeviceBuffer<PairRecord>    pairRecords; // child-pair records (pair_records.h) when the scene's leaves fit a link
    PairSceneDevice             pairsDev{};
    int                         traceKernel = 0; // as rf_renderer's: 2 selects the child-pair kernel
    bool                        usePairs() const { return pairsDev.records != nullptr && traceKernel == 2; }
};

extern "C" rf_status rf_traversal_scene_create(
    const rf_bvh_node*   nodes,
    std::uint64_t        numNodes,
    const rf_positions*  triangles,
    std::uint64_t        numTriangles,
    int32_t              device,
    rf_traversal_scene** out)
{
    if (!nodes || !triangles || !out || numTriangles == 0) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_traversal_scene_create: null or empty argument");
    bool      ordered = true;
    rf_status st = validateBvh(nodes, numNodes, numTriangles, ordered);
    if (st != RF_OK) return st;
    auto s = std::make_unique<rf_traversal_scene>();
    s->ordered = ordered;
    st = selectDevice(device, s->device, s->numSms);
    if (st != RF_OK) return st;
    DeviceBuffer<rf_bvh_node> rawNodes;
    DeviceBuffer<float>       rawTris;
    RF_CUDA(rawNodes.allocate(numNodes));
    RF_CUDA(rawTris.allocate(numTriangles * 9));
    RF_CUDA(cudaMemcpy(rawNodes.ptr, nodes, numNodes * sizeof(rf_bvh_node), cudaMemcpyHostToDevice));
    RF_CUDA(cudaMemcpy(rawTris.ptr, triangles, numTriangles * sizeof(rf_positions), cudaMemcpyHostToDevice));
    RF_CUDA(s->nodes.allocate(numNodes));
    RF_CUDA(s->cursor.allocate(1));
    RF_CUDA(s->tris.allocate(TRI_STRIDE * numTriangles));
    k_pack_nodes<<<s->numSms * 4, 256>>>(rawNodes.ptr, numNodes, s->nodes.ptr);
    k_pack_triangles<<<s->numSms * 4, 256>>>(rawTris.ptr, 3, numTriangles, s->tris.ptr);
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaDeviceSynchronize());
    s->numNodes = numNodes, s->numTris = numTriangles;
    const PairScene ps = buildPairRecords(nodes, numNodes);
    if (ps.usable)
    {
        RF_CUDA(s->pairRecords.allocate(ps.records.size() + 1));
        if (!ps.records.empty())
            RF_CUDA(cudaMemcpy(s->pairRecords.ptr, ps.records.data(), ps.records.size() * sizeof(PairRecord), cudaMemcpyHostToDevice));
        s->pairsDev.records = s->pairRecords.ptr;
        std::memcpy(s->pairsDev.rootBox, ps.rootBox, sizeof(ps.rootBox));
        s->pairsDev.rootLink = ps.rootLink;
    }
    *out = s.release();
    return RF_OK;
}

extern "C" rf_status rf_traversal_scene_set_kernel(rf_traversal_scene* s, std::int32_t traceKernel)
{
    if (!s || traceKernel < 0 || traceKernel > 2) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_traversal_scene_set_kernel: bad argument");
    s->traceKernel = traceKernel;
    return RF_OK;
}

extern "C" void rf_traversal_scene_destroy(rf_traversal_scene* s)
{
    if (s) cudaSetDevice(s->device);
    delete s;
}

extern "C" rf_status rf_ray_intersect_bvh(
    rf_traversal_scene* s,
    const float*        rays,
    std::uint64_t       numRays,
    float               rayTMax,
    std::uint8_t*       outHit,
    float*              outPT,
    std::uint32_t*      outNodes)
{
    if (!s || (!rays && numRays)) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_ray_intersect_bvh: null argument");
    if (numRays == 0) return RF_OK;
    RF_CUDA(cudaSetDevice(s->device));
    DeviceBuffer<float>         dRays;
    DeviceBuffer<std::uint8_t>  dHit;
    DeviceBuffer<float4>        dPT;
    DeviceBuffer<std::uint32_t> dNodes;
    RF_CUDA(dRays.allocate(numRays * 6));
    RF_CUDA(cudaMemcpy(dRays.ptr, rays, numRays * 6 * sizeof(float), cudaMemcpyHostToDevice));
    if (outHit) RF_CUDA(dHit.allocate(numRays));
    if (outPT) RF_CUDA(dPT.allocate(numRays));
    if (outNodes) RF_CUDA(dNodes.allocate(numRays));
    if (numRays >= 0xFFFFFFF0ull) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_ray_intersect_bvh: at most 2^32-16 rays per call");
    const std::uint64_t blocksNeeded = (numRays + TRACE_BLOCK_THREADS - 1) / TRACE_BLOCK_THREADS;
    const int           grid = static_cast<int>(std::min<std::uint64_t>(blocksNeeded, static_cast<std::uint64_t>(s->numSms) * 4));
    RF_CUDA(cudaMemset(s->cursor.ptr, 0, sizeof(std::uint32_t)));
    if (s->usePairs())
        k_intersect_batch_pairs<<<grid, TRACE_BLOCK_THREADS>>>(
            s->pairsDev, s->tris.ptr, s->ordered, s->tuning, dRays.ptr, static_cast<std::uint32_t>(numRays), rayTMax, s->cursor.ptr, dHit.ptr, dPT.ptr,
            dNodes.ptr);
    else
        k_intersect_batch<<<grid, TRACE_BLOCK_THREADS>>>(
            s->nodes.ptr, s->tris.ptr, s->ordered, s->tuning, dRays.ptr, static_cast<std::uint32_t>(numRays), rayTMax, s->cursor.ptr,
            dHit.ptr, dPT.ptr, dNodes.ptr);
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaDeviceSynchronize());
    if (outHit) RF_CUDA(cudaMemcpy(outHit, dHit.ptr, numRays, cudaMemcpyDeviceToHost));
    if (outPT) RF_CUDA(cudaMemcpy(outPT, dPT.ptr, numRays * sizeof(float4), cudaMemcpyDeviceToHost));
    if (outNodes) RF_CUDA(cudaMemcpy(outNodes, dNodes.ptr, numRays * sizeof(std::uint32_t), cudaMemcpyDeviceToHost));
    return RF_OK;
}

// Click-to-focus, pt/main.cpp:198-226: the ray through the cursor, rayIntersectBvh(ray, bvhNodes, positions, 1000.f,
// hitData), focusDistance = dot(hitData.p - cameraPosition, cameraForward).
extern "C" rf_status rf_pick_focus_distance(
    rf_traversal_scene* s,
    const rf_camera*    camera,
    const float         cameraPosition[3],
    const float         cameraForward[3],
    const double        cursorX,
    const double        cursorY,
    const std::int32_t  windowWidth,
    const std::int32_t  windowHeight,
    std::uint8_t*       outHit,
    float*              outFocusDistance)
{
    if (!s || !camera || !cameraPosition || !cameraForward || !outHit || !outFocusDistance || windowWidth <= 0 || windowHeight <= 0)
        return setError(RF_ERROR_INVALID_ARGUMENT, "rf_pick_focus_distance: bad argument");
    *outHit = 0;
    // main.cpp:207-208: clicks outside the window are ignored
    if (!(cursorX >= 0.0 && cursorX < static_cast<double>(windowWidth) && cursorY >= 0.0 && cursorY < static_cast<double>(windowHeight))) return RF_OK;
    const float u = static_cast<float>(cursorX) / static_cast<float>(windowWidth);
    const float v = 1.f - static_cast<float>(cursorY) / static_cast<float>(windowHeight);
    // generateCameraRay, camera.cpp:44-51
    const V3    origin = v3(camera->origin);
    const V3    dir = normalize(((v3(camera->lower_left_corner) + v3(camera->horizontal) * u) + v3(camera->vertical) * v) - origin);
    const float ray[6] = {origin.x, origin.y, origin.z, dir.x, dir.y, dir.z};
    float       pt[4] = {};
    const rf_status st = rf_ray_intersect_bvh(s, ray, 1, 1000.f, outHit, pt, nullptr);
    if (st != RF_OK) return st;
    if (*outHit) *outFocusDistance = dot(v3(pt) - v3(cameraPosition), v3(cameraForward));
    return RF_OK;
}

extern "C" rf_status rf_bvh_visualizer_node_counts(
    rf_traversal_scene* s,
    const rf_camera*    camera,
    std::uint32_t       width,
    std::uint32_t       height,
    float               rayTMax,
    std::uint32_t*      outNodes,
    float*              deviceMs)
{
    if (!s || !camera || !outNodes || width == 0 || height == 0) return setError(RF_ERROR_INVALID_ARGUMENT, "rf_bvh_visualizer_node_counts: bad argument");
    RF_CUDA(cudaSetDevice(s->device));
    const std::uint64_t         numPixels = static_cast<std::uint64_t>(width) * height;
    DeviceBuffer<std::uint32_t> dNodes;
    RF_CUDA(dNodes.allocate(numPixels));
    cudaEvent_t e0, e1;
    RF_CUDA(cudaEventCreate(&e0));
    RF_CUDA(cudaEventCreate(&e1));
    RF_CUDA(cudaMemset(s->cursor.ptr, 0, sizeof(std::uint32_t)));
    RF_CUDA(cudaEventRecord(e0));
    if (s->usePairs())
        k_visualizer_pairs<<<s->numSms * 4, TRACE_BLOCK_THREADS>>>(
            s->pairsDev, s->tris.ptr, s->ordered, s->tuning, *camera, width, height, rayTMax, s->cursor.ptr, dNodes.ptr);
    else
        k_visualizer<<<s->numSms * 4, TRACE_BLOCK_THREADS>>>(
            s->nodes.ptr, s->tris.ptr, s->ordered, s->tuning, *camera, width, height, rayTMax, s->cursor.ptr, dNodes.ptr);
    RF_CUDA(cudaEventRecord(e1));
    RF_CUDA(cudaGetLastError());
    RF_CUDA(cudaDeviceSynchronize());
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (deviceMs) *deviceMs = ms;
    RF_CUDA(cudaMemcpy(outNodes, dNodes.ptr, numPixels * sizeof(std::uint32_t), cudaMemcpyDeviceToHost));
    return RF_OK;
}

extern "C" int32_t     rf_has_cuda_kernels(void) { return 1; }
#define RF_STR2(x) #x
#define RF_STR(x) RF_STR2(x)
extern "C" const char* rf_build_info(void) { return "rayfinder_b200 sm_100a -fmad=false cudart " RF_STR(CUDART_VERSION); }

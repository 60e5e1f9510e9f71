/*
 * rayfinder_b200 — C-ABI of the B200-native render path.
 *
 * This header is the drop-in boundary for the reference's render path (Nelarius/rayfinder @ 634128c).
 * The reference has no FFI layer: the path sits behind one C++ class (nlrs::ReferencePathTracer) and
 * one free function (nlrs::rayIntersectBvh).  Every entry point below names the reference interface
 * it replaces (paths relative to the reference's src/).  All structs are plain C PODs whose byte
 * layout equals the reference struct they mirror, so a maintainer can reinterpret_cast spans of the
 * reference's own containers.
 *
 * Conventions: every function returns rf_status (0 = ok) unless it is a pure getter; on failure
 * rf_last_error() returns a thread-local, NUL-terminated message (the reference throws
 * std::runtime_error with the same text where one exists).  Handles are opaque and, like the
 * reference objects, not thread-safe.  There is NO CPU fallback: creating a renderer or a traversal
 * scene without a CUDA device fails with RF_ERROR_CUDA.
 */
#ifndef RAYFINDER_B200_H
#define RAYFINDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t rf_status;
enum
{
    RF_OK = 0,
    RF_ERROR_INVALID_ARGUMENT = 1,
    RF_ERROR_CUDA = 2,
    RF_ERROR_IO = 3,
    RF_ERROR_FORMAT = 4,       /* .pt magic/version errors (pt-format/pt_format.cpp:277-291) */
    RF_ERROR_OUT_OF_RANGE = 5, /* sky params out of range (hw-skymodel/hw_skymodel.c:148-161) */
};

const char* rf_last_error(void);

/* ---- POD mirrors of the reference's scene structs --------------------------------------------- */

/* nlrs::BvhNode, common/bvh.hpp:13-21 (48 B; Aabb = common/aabb.hpp:12-27). */
typedef struct rf_bvh_node
{
    float    aabb_min[3];
    float    pad0;
    float    aabb_max[3];
    float    pad1;
    uint32_t triangles_offset;
    uint32_t second_child_offset;
    uint32_t triangle_count;
    uint32_t split_axis;
} rf_bvh_node;

/* nlrs::Positions, common/triangle_attributes.hpp:7-12 (36 B, CPU-side triangle). */
typedef struct rf_positions
{
    float v0[3];
    float v1[3];
    float v2[3];
} rf_positions;

/* nlrs::PositionAttribute, pt-format/vertex_attributes.hpp:7-15 (48 B). */
typedef struct rf_position_attribute
{
    float p0[3];
    float pad0;
    float p1[3];
    float pad1;
    float p2[3];
    float pad2;
} rf_position_attribute;

/* nlrs::VertexAttributes, pt-format/vertex_attributes.hpp:17-35 (80 B). */
typedef struct rf_vertex_attributes
{
    float    n0[3];
    float    pad0;
    float    n1[3];
    float    pad1;
    float    n2[3];
    float    pad2;
    float    uv0[2];
    float    uv1[2];
    float    uv2[2];
    uint32_t texture_idx;
    uint32_t pad3;
} rf_vertex_attributes;

/* nlrs::Texture, common/texture.hpp:10-49: BGRA8 pixels (b | g<<8 | r<<16 | a<<24), row-major. */
typedef struct rf_texture
{
    const uint32_t* pixels;
    uint32_t        width;
    uint32_t        height;
} rf_texture;

/* nlrs::Scene, pt/reference_path_tracer.hpp:45-51: four non-owning spans.  Data is copied to the
 * device inside rf_renderer_create; the caller may free it afterwards (as pt/main.cpp:176-177 does). */
typedef struct rf_scene
{
    const rf_bvh_node*           bvh_nodes;
    uint64_t                     num_bvh_nodes;
    const rf_position_attribute* position_attributes;
    uint64_t                     num_position_attributes;
    const rf_vertex_attributes*  vertex_attributes;
    uint64_t                     num_vertex_attributes;
    const rf_texture*            base_color_textures;
    uint64_t                     num_base_color_textures;
} rf_scene;

/* nlrs::Camera, common/camera.hpp:10-21 (19 floats). */
typedef struct rf_camera
{
    float origin[3];
    float lower_left_corner[3];
    float horizontal[3];
    float vertical[3];
    float up[3];
    float right[3];
    float lens_radius;
} rf_camera;

/* nlrs::SamplingParams, pt/reference_path_tracer.hpp:26-32 (defaults 128 / 4). */
typedef struct rf_sampling_params
{
    uint32_t num_samples_per_pixel;
    uint32_t num_bounces;
} rf_sampling_params;

/* nlrs::Sky, pt/aligned_sky_state.hpp:15-23 (defaults 1, {1,1,1}, 30, 0). */
typedef struct rf_sky
{
    float turbidity;
    float albedo[3];
    float sun_zenith_degrees;
    float sun_azimuth_degrees;
} rf_sky;

/* nlrs::RenderParameters, pt/reference_path_tracer.hpp:34-43. */
typedef struct rf_render_parameters
{
    uint32_t           framebuffer_width;
    uint32_t           framebuffer_height;
    rf_camera          camera;
    rf_sampling_params sampling_params;
    rf_sky             sky;
    float              exposure;
} rf_render_parameters;

/* nlrs::RendererDescriptor, pt/reference_path_tracer.hpp:53-57. */
typedef struct rf_renderer_descriptor
{
    rf_render_parameters render_params;
    int32_t              max_framebuffer_width;
    int32_t              max_framebuffer_height;
} rf_renderer_descriptor;

/* nlrs::AlignedSkyState, pt/aligned_sky_state.hpp:34-41 (160 B, the block the WGSL reads). */
typedef struct rf_sky_state
{
    float params[27];
    float sky_radiances[3];
    float solar_radiances[3];
    float padding1[3];
    float sun_direction[3];
    float padding2;
} rf_sky_state;

/* Work counters of the frames rendered since the last rf_renderer_reset_stats (new: the reference
 * only records the render-pass duration, pt/reference_path_tracer.cpp:668-716).  A "ray" is one
 * call of the reference's rayIntersectBvh (wgsl:190) or shadowRay (wgsl:202). */
typedef struct rf_frame_stats
{
    uint64_t frames;
    uint64_t paths;                /* primary rays generated */
    uint64_t closest_rays;         /* rayIntersectBvh calls */
    uint64_t shadow_rays;          /* shadowRay calls */
    uint64_t closest_nodes_visited;
    uint64_t closest_triangles_tested;
    uint64_t shadow_nodes_visited;
    uint64_t shadow_triangles_tested;
    double   device_ms_total;      /* CUDA-event time of the render passes */
    double   device_ms_trace;      /* per-stage CUDA-event time, only filled when stage timing is on: the */
    double   device_ms_shade;      /* traversal launches (closest-hit + shadow rays), shading, the rest   */
    double   device_ms_other;
    uint64_t kernel_launches;      /* kernels launched by rf_renderer_render since the last reset */
    uint32_t sub_frames;           /* schedule in effect: tile sets per frame, */
    uint32_t evict_max;            /* tail hand-over threshold (rf_renderer_set_pipeline / _set_tail_policy) */
    uint64_t node_records_loaded;  /* BVH records the traversal kernels loaded: one 64-byte child-pair record per interior
                                    * node a ray entered (csrc/traversal_pairs.cuh), or one 32-byte node per visit with the
                                    * per-node kernel — the memory work behind the visits above */
    uint32_t trace_kernel;         /* traversal kernel in effect: 2 = child-pair records, 1 = one node per visit */
    uint32_t persistent_kernel;    /* 1: frames run as one persistent launch (csrc/mega.cuh), 0: one launch per stage */
} rf_frame_stats;

/* ---- the renderer: nlrs::ReferencePathTracer (pt/reference_path_tracer.hpp:59-102) ------------- */

typedef struct rf_renderer rf_renderer;

/* ReferencePathTracer(const RendererDescriptor&, const GpuContext&, Scene)
 * (reference_path_tracer.cpp:131-481).  `device` < 0 selects the current CUDA device; the
 * GpuContext argument has no analogue.  Fails with the reference's message when the texture data
 * exceeds the 1 GiB binding limit (reference_path_tracer.cpp:256-263, pt/gpu_limits.hpp:20-25). */
rf_status rf_renderer_create(
    const rf_renderer_descriptor* desc,
    const rf_scene*               scene,
    int32_t                       device,
    rf_renderer**                 out);

/* ~ReferencePathTracer (reference_path_tracer.cpp:548-554). */
void rf_renderer_destroy(rf_renderer* r);

/* setRenderParameters (reference_path_tracer.cpp:556-563): any change resets the accumulation. */
rf_status rf_renderer_set_render_parameters(rf_renderer* r, const rf_render_parameters* params);

/* render (reference_path_tracer.cpp:565-704): one sample per pixel is traced and added to the HDR
 * accumulation buffer when accumulated < numSamplesPerPixel; frameCount always advances.  The
 * swap-chain / ImGui arguments have no analogue; the tonemapped image is read with
 * rf_renderer_read_display.  Asynchronous on the renderer's stream. */
rf_status rf_renderer_render(rf_renderer* r);

/* averageRenderpassDurationMs (reference_path_tracer.cpp:706-716): mean of the last <= 30 passes. */
float rf_renderer_average_renderpass_duration_ms(rf_renderer* r);
/* renderProgressPercentage (reference_path_tracer.cpp:718-722). */
float rf_renderer_render_progress_percentage(const rf_renderer* r);

/* New (the reference never reads its image back): copy the accumulated HDR *sum* buffer
 * `imageBuffer: array<vec3f>` (wgsl:32; 16-byte stride, W*H*4 floats, row 0 = top) to host memory
 * and return the sample count it holds. Synchronises the stream. */
rf_status rf_renderer_read_hdr(rf_renderer* r, float* dst_rgba, uint64_t num_floats, uint32_t* accumulated);
/* The display transform of fsMain (wgsl:59-63): estimator = sum/acc, acesFilmic(exposure*x),
 * pow(1/2.2), packed BGRA8 (the reference's swap-chain format), W*H u32. */
rf_status rf_renderer_read_display(rf_renderer* r, uint32_t* dst_bgra8, uint64_t num_pixels);

/* Device-side access for multi-GPU reduction and zero-copy consumers. */
void*     rf_renderer_hdr_device_ptr(rf_renderer* r);       /* float4[max_w*max_h] */
rf_status rf_renderer_set_stream(rf_renderer* r, void* cuda_stream);
rf_status rf_renderer_synchronize(rf_renderer* r);

/* The reference's frame counter (mFrameCount, reference_path_tracer.cpp:580) is the "seed" of the
 * blue-noise/R2 sequence (wgsl:603-616); expose it so runs are reproducible. */
rf_status rf_renderer_set_frame_count(rf_renderer* r, uint32_t frame_count);
uint32_t  rf_renderer_frame_count(const rf_renderer* r);
uint32_t  rf_renderer_accumulated_sample_count(const rf_renderer* r);

/* Multi-GPU: this renderer traces only the 32x32-pixel tiles t=(tx,ty) with (tx+ty) % world == rank
 * and leaves every other pixel of the HDR buffer exactly 0, so a sum-reduce over ranks is
 * bit-identical to a single-GPU frame.  Resets the accumulation. */
rf_status rf_renderer_set_tile_partition(rf_renderer* r, uint32_t rank, uint32_t world);

/* The exchange fused into the accumulation kernel (one process per GPU, same node).  The root rank allocates and exports an
 * exchange buffer of TWO full frames (64-byte CUDA IPC handle); every other rank maps it, and from then on the accumulation
 * kernel of every rank — the root included — also stores each owned pixel's accumulated value into half (frame & 1) of that
 * buffer, over NVLink peer memory on the other ranks.  A pixel has exactly one owner, so after all ranks have finished the
 * frame (any stream-ordered barrier, e.g. a 4-byte all-reduce) that half holds the full frame, bit-identical to the
 * single-GPU one, without moving W x H x 16 bytes through a reduction; rf_renderer_read_hdr / _read_display /
 * rf_renderer_exchange_device_ptr on the root then present it.  The double buffer is what makes the root's read safe: the
 * other ranks may already be storing frame N + 1 (into the other half) while the root still reads frame N, and they come
 * back to this half only after the barrier of frame N + 1, which the root enters — in stream order — after its read.
 * Contract: every rank is created with the same maxFramebufferSize, calls rf_renderer_render the same number of times
 * with the same parameters, and runs the barrier after every frame on the rendering stream, as
 * rayfinder_b200/distributed.py does.  The local HDR buffer (rf_renderer_hdr_device_ptr) is never written by another rank.
 * NULL detaches. */
rf_status rf_renderer_hdr_ipc_handle(rf_renderer* r, void* out_handle_64_bytes);
rf_status rf_renderer_set_hdr_peer(rf_renderer* r, const void* handle_64_bytes);
/* Root of an exchange: the half of the exchange buffer that holds the last complete frame (float4[W*H]); else NULL. */
void*     rf_renderer_exchange_device_ptr(rf_renderer* r);

rf_status rf_renderer_get_stats(rf_renderer* r, rf_frame_stats* out);
rf_status rf_renderer_reset_stats(rf_renderer* r);
/* Per-stage CUDA-event timing (adds event records between stages; off by default). */
rf_status rf_renderer_set_stage_timing(rf_renderer* r, int32_t enabled);
/* Scheduling knobs of the persistent traversal kernels (results never depend on them): a triangle round
 * runs once `tri_min` lanes of a warp have a triangle pending, idle lanes are refilled once `refill_min`
 * are idle, `blocks_per_sm` persistent 256-thread blocks are launched per SM (automatic until set: 4, or 2 per tile set
 * for a small frame, see rf_renderer_set_pipeline).  0 keeps a value. */
rf_status rf_renderer_set_tuning(rf_renderer* r, uint32_t tri_min, uint32_t refill_min, uint32_t blocks_per_sm);
/* How a frame is scheduled (results never depend on it).
 * persistent_kernel (0 / 1 / 2, -1 keeps; 2 = automatic is the default): 1 = the whole frame of a tile set is ONE persistent
 *   launch (csrc/mega.cuh): every block generates its own primary rays and keeps its paths to itself — ready rays of any
 *   bounce wait in shared-memory rings, 15 (or 7) warps trace, one warp shades 32 hits at a time, a path's shadow ray and
 *   next closest-hit ray run on the same lane, paths that lag behind are served first, and at the end of the frame a warp
 *   walks its last ray with all 32 lanes.  Nothing waits for a bounce to finish, which is what a small frame needs: the
 *   automatic schedule uses it when this GPU owns at most ~1.2 M pixels (a 1080p frame split over 2-8 GPUs) and paths have at
 *   least 6 bounces (3 between 0.15 M and 0.6 M pixels), and the staged pipeline (one launch per stage: raygen, trace, shade,
 *   ...) otherwise.
 * sub_frames (1..4, 0 keeps, -1 automatic = the default): the frame is traced as that many independent tile sets on separate
 *   CUDA streams.  Automatic: 1 with the persistent kernel; 2 with the staged pipeline, so that one set's traversal tail
 *   overlaps the other's work — except between ~1.2 M and ~3 M owned pixels (option "walk_in_place" on), where one set of
 *   full-size launches measured faster.
 * variant (0..15, -1 keeps) / block_threads (64, 128, 256, 0 keeps): compile-time scheduling variant and block size of the
 *   staged traversal kernel (the persistent kernel's block size is the option "mega_block"). */
rf_status rf_renderer_set_pipeline(rf_renderer* r, int32_t sub_frames, int32_t persistent_kernel, int32_t variant, int32_t block_threads);
/* Tail policy of the traversal launches (results never depend on it).  Once a launch's ray queue is dry, a warp left
 * with <= evict_max rays keeps them for four more loop rounds (most of them are short and end there), then writes the
 * traversal state of the rest to a device buffer and exits; a follow-up launch gives each of those rays a whole warp
 * (32 consecutive nodes loaded and slab-tested per memory round trip, csrc/straggler.cuh), which walks the long ones
 * ~2.5x faster than a lone lane can.  0 = off (every ray ends on the lane it started on), -1 = automatic (the default: off while
 * the option "walk_in_place" is on — a warp then walks its last ray in place, which measured faster than handing it over —, else
 * 8 when this GPU owns at most ~0.6 M pixels), up to 32. */
rf_status rf_renderer_set_tail_policy(rf_renderer* r, int32_t evict_max);
/* Named scheduling / debugging knobs (results never depend on them; unknown names are an error):
 *   "shade_wait"   persistent kernel: 0.5 us naps its shading warp takes to let a batch of 32 hits fill (default 16)
 *   "mega_block"   persistent kernel: threads per block, 512 (2 blocks per SM, 15 traversal warps + 1 shading warp each; the
 *                  default) or 256 (4 per SM, 7 + 1)
 *   "mega_slots"   persistent kernel: path slots per block = paths a block keeps in flight (a multiple of 32, at most four times
 *                  the block size; 0 = automatic: the block's share of the pixels, so that a small frame's paths all start at
 *                  once, or that share split into equal waves when it exceeds the maximum)
 *   "walk_in_place" staged pipeline: 1 (default) = once a traversal launch's queue is dry, a warp left with ONE ray walks it with
 *                  all 32 lanes through 32-node windows, in place (csrc/straggler.cuh traceWarpRay); 0 = the ray ends on its lane
 *   "priority_mode" persistent kernel: 0 = paths that lag behind the block's other paths are served first (the default), 1 = plain
 *                  FIFO, 2 = lagging paths first only once the block has taken its last pixels
 *   "tail_paths"   persistent kernel: a block that has taken its last pixels and has at most n live paths left gives each ray a
 *                  whole warp (value n + 1; 1 = never; 0 = automatic: two per traversal warp)
 *   "evict_delay"  loop rounds a warp keeps its last rays before handing them to the tail launch (default 4)
 *   "trace_stack"  force at least this many traversal-stack entries (<= 32; 0 = what the scene needs, the default)
 *   "trace_kernel" 0 / 1: traversal over 32-byte nodes, one per visit (csrc/traversal.cuh; the default); 2: over 64-byte
 *                  child-pair records, one per interior node entered (csrc/traversal_pairs.cuh)
 *   "pair_variant" scheduling variant of the child-pair kernel: bits 0-1 = rounds per warp vote - 1, bit 2 = closest-hit rays
 *                  do not push far children that miss whatever tmax is (1, 3, 5 or 7)
 *   "tail_window_mode" how the warp-per-ray tail kernel fetches its 32-node windows: 0 = one LDG.256 per lane; 1 = one
 *                  1 KB bulk asynchronous copy (cp.async.bulk + mbarrier) into shared memory, with the window of the newest
 *                  out-of-window stack entry copied ahead of time while the current window is walked (csrc/straggler.cuh)
 *   "stage_debug"  1: print the per-launch spans of every stage-timed frame to stderr */
rf_status rf_renderer_set_option(rf_renderer* r, const char* name, int64_t value);

/* ---- the deferred renderer's lighting + resolve passes as a second integrator -------
 * pt/deferred_renderer_lighting_pass.wgsl:96-186 and pt/deferred_renderer_resolve_pass.wgsl:34-53, i.e. the
 * `lightingPass` and `resolvePass` of DeferredRenderer::render (pt/deferred_renderer.cpp:340-375).  The G-buffer
 * the reference's rasteriser produces is an input here. */
typedef struct rf_deferred_lighting_params
{
    float    inverse_view_reverse_z_projection[16]; /* Uniforms.inverseViewReverseZProjectionMat, column-major */
    float    camera_eye[4];                          /* Uniforms.cameraEye */
    uint32_t framebuffer_width;                      /* Uniforms.framebufferSize */
    uint32_t framebuffer_height;
    uint32_t frame_count;                            /* Uniforms.frameCount (0 restarts the moving average) */
    float    exposure;                               /* resolve pass Uniforms.exposure */
    rf_sky   sky;                                    /* -> SkyState, as for the path tracer */
} rf_deferred_lighting_params;

/* One frame of the lighting pass (sun sample at the G-buffer surface, one bounce, sun sample or sky there) followed by the
 * resolve pass's moving average.  G-buffer, host memory, row-major, row 0 = top, what textureLoad returns per texel:
 * albedo and normal 4 floats (rgb + unused; the normal ENCODED as 0.5 n + 0.5), depth 1 float (reverse Z: 0 = no
 * surface).  The framebuffer size must not exceed the renderer's maxFramebufferSize; the scene is the renderer's. */
rf_status rf_renderer_render_deferred_lighting(
    rf_renderer*                       r,
    const rf_deferred_lighting_params* params,
    const float*                       gbuffer_albedo,
    const float*                       gbuffer_normal,
    const float*                       gbuffer_depth);
/* sampleBuffer and accumulationBuffer (3 floats per pixel) and the resolve pass's BGRA8 output of the last deferred
 * frame; each pointer may be NULL.  The sample buffer shares the renderer's per-frame radiance scratch: read it before the
 * next rf_renderer_render / rf_renderer_render_deferred_lighting call (the accumulation buffer and the display persist). */
rf_status rf_renderer_read_deferred(rf_renderer* r, float* sample_rgb, float* accumulation_rgb, uint32_t* display_bgra8);

/* ---- the CPU traversal twin: nlrs::rayIntersectBvh (common/ray_intersection.hpp:43-49) --------- */

typedef struct rf_traversal_scene rf_traversal_scene;

/* Uploads (bvhNodes, triangles) as passed to rayIntersectBvh by bvh-visualizer/main.cpp:71,
 * pt/main.cpp:217 and tests/bvh.cpp:92. */
rf_status rf_traversal_scene_create(
    const rf_bvh_node*   bvh_nodes,
    uint64_t             num_bvh_nodes,
    const rf_positions*  triangles,
    uint64_t             num_triangles,
    int32_t              device,
    rf_traversal_scene** out);
void rf_traversal_scene_destroy(rf_traversal_scene* s);
/* Traversal kernel of the calls below (results never depend on it): 0 / 1 = one 32-byte node per visit (the default),
 * 2 = one 64-byte child-pair record per interior node entered (csrc/traversal_pairs.cuh; scenes whose leaves do not fit
 * its links keep the per-node kernel).  The renderer's counterpart is rf_renderer_set_option("trace_kernel", ...). */
rf_status rf_traversal_scene_set_kernel(rf_traversal_scene* s, int32_t trace_kernel);

/* Batched rayIntersectBvh (ray_intersection.cpp:138-213) on host buffers.  rays: n x 6 floats
 * (origin, direction).  Outputs (each may be NULL): out_hit n x u8 (the bool result); out_p_t n x 4
 * floats = Intersection{p, t} (ray_intersection.hpp:15-19; zeros on miss); out_nodes_visited n x u32 =
 * BvhStats::nodesVisited (:37-40). */
rf_status rf_ray_intersect_bvh(
    rf_traversal_scene* s,
    const float*        rays,
    uint64_t            num_rays,
    float               ray_t_max,
    uint8_t*            out_hit,
    float*              out_p_t,
    uint32_t*           out_nodes_visited);

/* The pixel loop of bvh-visualizer/main.cpp:60-78: u = j/W, v = 1-(i+1)/H, generateCameraRay
 * (camera.cpp:44-51), rayIntersectBvh(..., ray_t_max, ...); out_nodes_visited is W*H u32, row-major,
 * row 0 = top.  `device_ms` (may be NULL) receives the kernel time. */
rf_status rf_bvh_visualizer_node_counts(
    rf_traversal_scene* s,
    const rf_camera*    camera,
    uint32_t            width,
    uint32_t            height,
    float               ray_t_max,
    uint32_t*           out_nodes_visited,
    float*              device_ms);

/* Click-to-focus (pt/main.cpp:198-226): the camera ray through the cursor (u = x / width, v = 1 - y / height,
 * generateCameraRay), rayIntersectBvh with rayTMax 1000, and on a hit focusDistance = dot(hit.p - camera_position,
 * camera_forward).  A cursor outside the window or a miss leaves *out_focus_distance untouched and *out_hit 0, as
 * the reference leaves its controller untouched. */
rf_status rf_pick_focus_distance(
    rf_traversal_scene* s,
    const rf_camera*    camera,
    const float         camera_position[3],
    const float         camera_forward[3],
    double              cursor_x,
    double              cursor_y,
    int32_t             window_width,
    int32_t             window_height,
    uint8_t*            out_hit,
    float*              out_focus_distance);

/* ---- host-side pieces the path keeps (no GPU needed) ------------------------------------------- */

/* createCamera (common/camera.cpp:7-42). vfov in radians (Angle::asRadians). */
rf_status rf_create_camera(
    const float origin[3],
    const float look_at[3],
    float       aperture,
    float       focus_distance,
    float       vfov_radians,
    float       aspect_ratio,
    rf_camera*  out);

/* AlignedSkyState(const Sky&) (pt/aligned_sky_state.hpp:43-70) on top of sky_state_new
 * (hw-skymodel/hw_skymodel.c:141-180).  Returns RF_ERROR_OUT_OF_RANGE (with the failing field named
 * in rf_last_error) where sky_state_new returns a non-success sky_state_result. */
rf_status rf_sky_state_new(const rf_sky* sky, rf_sky_state* out);

/* buildBvh (common/bvh.cpp:263-291).  out_nodes needs room for 2*n-1 nodes; out_triangle_indices
 * (n x u64) is Bvh::triangleIndices (old index -> new index, common/bvh.hpp:24-28). */
rf_status rf_build_bvh(
    const rf_positions* triangles,
    uint64_t            num_triangles,
    rf_bvh_node*        out_nodes,
    uint64_t*           out_num_nodes,
    uint64_t*           out_triangle_indices);

/* buildBvh on the GPU (csrc/bvh_build.cu): same arguments and byte-identical results as
 * rf_build_bvh — nodes, padding words and triangleIndices — built level by level with data-parallel kernels
 * (atomic box reductions, binned SAH sweep per node, std::partition reproduced with one scan).  `device` < 0 selects
 * the current CUDA device; `out_device_ms` (may be NULL) receives the device time of the build without the
 * host<->device copies.  Fails with RF_ERROR_CUDA when there is no CUDA device (rf_build_bvh is the host builder). */
rf_status rf_build_bvh_device(
    const rf_positions* triangles,
    uint64_t            num_triangles,
    int32_t             device,
    rf_bvh_node*        out_nodes,
    uint64_t*           out_num_nodes,
    uint64_t*           out_triangle_indices,
    float*              out_device_ms);

/* 0 (default): the build is one persistent launch with grid-wide barriers between its phases; 1: one launch per phase and
 * level (round 1's path, kept for A/B timing).  Same bytes either way.  Process-wide. */
void rf_build_bvh_device_set_mode(int32_t level_kernels);
/* Diagnostics of the last single-launch build: milliseconds its first block spent in each phase, summed over the levels
 * (12 floats: boxes, decide, buckets, sweep, scan, -, pair, permute, -, leaf scan + node records, -, block-local subtrees);
 * returns the number of grid-wide levels. */
uint32_t rf_build_bvh_device_last_phases(float* out_phase_ms);
/* rf_build_bvh_device keeps its device arrays (~0.5 KB per triangle) between calls; this frees them. */
void rf_build_bvh_device_release(void);

/* ---- .pt container: nlrs::PtFormat + serialize/deserialize (pt-format/pt_format.hpp:18-43) ----- */

typedef struct rf_pt_file rf_pt_file;

enum
{
    RF_PT_BVH_NODES = 0,                      /* 48 B  */
    RF_PT_BVH_POSITION_ATTRIBUTES = 1,        /* 36 B  nlrs::Positions */
    RF_PT_TRIANGLE_POSITION_ATTRIBUTES = 2,   /* 48 B  */
    RF_PT_TRIANGLE_VERTEX_ATTRIBUTES = 3,     /* 80 B  */
    RF_PT_VERTEX_POSITIONS = 4,               /* 16 B vec4 */
    RF_PT_VERTEX_NORMALS = 5,                 /* 16 B vec4 */
    RF_PT_VERTEX_TEX_COORDS = 6,              /* 8 B vec2 */
    RF_PT_VERTEX_INDICES = 7,                 /* 4 B */
    RF_PT_MODEL_VERTEX_POSITIONS = 8,         /* slices: 16 B {u64 offsetIdx, u64 numElements} */
    RF_PT_MODEL_VERTEX_NORMALS = 9,
    RF_PT_MODEL_VERTEX_TEX_COORDS = 10,
    RF_PT_MODEL_VERTEX_INDICES = 11,
    RF_PT_MODEL_BASE_COLOR_TEXTURE_INDICES = 12, /* 4 B */
    RF_PT_NUM_ARRAYS = 13,
};

rf_status rf_pt_create(rf_pt_file** out);  /* PtFormat() = default */
void      rf_pt_destroy(rf_pt_file* f);
/* deserialize(InputStream&, PtFormat&) (pt_format.cpp:271-321) from a file / a memory buffer. */
rf_status rf_pt_load(const char* path, rf_pt_file** out);
rf_status rf_pt_load_memory(const void* data, uint64_t size, rf_pt_file** out);
/* serialize(OutputStream&, const PtFormat&) (pt_format.cpp:240-269). */
rf_status rf_pt_save(const rf_pt_file* f, const char* path);
rf_status rf_pt_save_memory(const rf_pt_file* f, void* dst, uint64_t capacity, uint64_t* size);
/* Array access: element count, element size and a pointer valid until the file is modified. */
rf_status rf_pt_array(const rf_pt_file* f, int32_t which, const void** data, uint64_t* count, uint64_t* elem_size);
rf_status rf_pt_set_array(rf_pt_file* f, int32_t which, const void* data, uint64_t count);
uint64_t  rf_pt_num_textures(const rf_pt_file* f);
rf_status rf_pt_texture(const rf_pt_file* f, uint64_t idx, rf_texture* out);
rf_status rf_pt_add_texture(rf_pt_file* f, const uint32_t* pixels, uint32_t width, uint32_t height);
/* Fill an rf_scene (the four spans of pt/main.cpp:150-155) pointing into `f`; `textures` must have
 * room for rf_pt_num_textures(f) entries. */
rf_status rf_pt_scene(const rf_pt_file* f, rf_scene* out, rf_texture* textures);

/* ---- scene baking: nlrs::PtFormat(gltfPath) (pt-format/pt_format.cpp:20-151), the step before the path ---------- */

/* Texture::fromMemory (common/texture.cpp:12-52): decode a PNG or baseline JPEG as stbi_load_from_memory(..., 4) does and
 * pack b | g << 8 | r << 16 | 255 << 24.  out_bgra == NULL queries the size; otherwise it needs width * height entries. */
rf_status rf_texture_from_memory(const void* data, uint64_t size, uint32_t* out_bgra, uint64_t capacity_pixels, uint32_t* width, uint32_t* height);

/* PtFormat(std::filesystem::path gltfPath) (pt-format/pt_format.cpp:20-151): GltfModel (common/gltf_model.cpp:266-465) ->
 * FlattenedModel (common/flattened_model.cpp:8-46) -> buildBvh + reorderAttributes -> the arrays serialize() writes.  Reads
 * .glb and .gltf (external or base64 buffers and images); host only.  Errors carry the reference's messages ("The gltf file
 * {} does not exist.", "Failed to parse gltf file {}.", "Failed to load gltf buffers for {}.", "The image {} does not
 * exist.").  pt-format-tool (pt-format-tool/main.cpp:14-35) = rf_bake_gltf + rf_pt_save. */
rf_status rf_bake_gltf(const char* gltf_path, rf_pt_file** out);

/* Introspection used by the tests / bench: 1 when built with CUDA kernels for sm_100a. */
int32_t     rf_has_cuda_kernels(void);
const char* rf_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* RAYFINDER_B200_H */

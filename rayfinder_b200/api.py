"""Host-side mirror of the reference's interface for the render path, on top of the C-ABI.

Names, argument meaning and error behaviour follow the reference (paths relative to its src/):

=============================  ===================================================================
here                           reference
=============================  ===================================================================
``create_camera``              ``createCamera``                  common/camera.cpp:7-42
``fly_camera``                 ``FlyCameraController::getCamera``  pt/fly_camera_controller.cpp:12-22
``Sky`` / ``sky_state``        ``Sky`` / ``AlignedSkyState``     pt/aligned_sky_state.hpp:15-71
``build_bvh``                  ``buildBvh``                      common/bvh.cpp:263-291
``reorder_attributes``         ``reorderAttributes``             common/bvh.hpp:36-46
``PtFormat``                   ``PtFormat`` + (de)serialize      pt-format/pt_format.{hpp,cpp}
``ReferencePathTracer``        ``ReferencePathTracer``           pt/reference_path_tracer.{hpp,cpp}
``TraversalScene``             ``rayIntersectBvh`` + the bvh-visualizer loop
                                                                 common/ray_intersection.cpp:138-213,
                                                                 bvh-visualizer/main.cpp:60-78
=============================  ===================================================================

Exceptions replace the reference's ``std::runtime_error`` (same messages where the reference has one).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import capi
from .capi import RayfinderError, check, lib

# numpy views of the reference PODs -----------------------------------------------------------------
BVH_NODE_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("pad0", "<f4"), ("aabb_max", "<f4", 3), ("pad1", "<f4"),
                           ("triangles_offset", "<u4"), ("second_child_offset", "<u4"),
                           ("triangle_count", "<u4"), ("split_axis", "<u4")])
POSITIONS_DTYPE = np.dtype([("v0", "<f4", 3), ("v1", "<f4", 3), ("v2", "<f4", 3)])
POSITION_ATTRIBUTE_DTYPE = np.dtype([("p0", "<f4", 3), ("pad0", "<f4"), ("p1", "<f4", 3), ("pad1", "<f4"),
                                     ("p2", "<f4", 3), ("pad2", "<f4")])
VERTEX_ATTRIBUTES_DTYPE = np.dtype([("n0", "<f4", 3), ("pad0", "<f4"), ("n1", "<f4", 3), ("pad1", "<f4"),
                                    ("n2", "<f4", 3), ("pad2", "<f4"), ("uv0", "<f4", 2), ("uv1", "<f4", 2),
                                    ("uv2", "<f4", 2), ("texture_idx", "<u4"), ("pad3", "<u4")])
assert BVH_NODE_DTYPE.itemsize == 48 and POSITIONS_DTYPE.itemsize == 36
assert POSITION_ATTRIBUTE_DTYPE.itemsize == 48 and VERTEX_ATTRIBUTES_DTYPE.itemsize == 80

FLT_MAX = float(np.finfo(np.float32).max)


def _f3(v) -> C.Array:
    return (C.c_float * 3)(*[float(x) for x in v])


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


# camera ---------------------------------------------------------------------------------------------
def degrees_to_radians(deg: float) -> float:
    """``Angle::degrees`` (common/units/angle.hpp:12-15): fp32 ``deg * pi_v<float> / 180.0f``."""
    return float(np.float32(np.float32(deg) * np.float32(math.pi)) / np.float32(180.0))


def create_camera(origin, look_at, aperture: float, focus_distance: float, vfov_radians: float,
                  aspect_ratio: float) -> capi.Camera:
    cam = capi.Camera()
    check(lib().rf_create_camera(_f3(origin), _f3(look_at), aperture, focus_distance, vfov_radians,
                                 aspect_ratio, C.byref(cam)))
    return cam


def fly_camera(width: int, height: int, position=(1.22, 1.25, -1.25), yaw_degrees: float = 129.64,
               pitch_degrees: float = -13.73, vfov_degrees: float = 70.0, aperture: float = 0.0,
               focus_distance: float = 10.0) -> capi.Camera:
    """The reference's default interactive view (pt/fly_camera_controller.hpp:47-52 with the UI's 70 degree
    vfov, pt/main.cpp:49,314): ``cameraOrientation`` (fly_camera_controller.cpp:138-148) + ``getCamera``."""
    f32 = np.float32
    yaw, pitch = f32(degrees_to_radians(yaw_degrees)), f32(degrees_to_radians(pitch_degrees))
    # cosf/sinf of the fp32 angles (evaluated in double, rounded once to fp32)
    cy, sy, cp, sp = (f32(fn(float(a))) for fn, a in ((math.cos, yaw), (math.sin, yaw), (math.cos, pitch), (math.sin, pitch)))
    fwd = np.array([cy * cp, sp, sy * cp], dtype=f32)
    d = f32(f32(fwd[0] * fwd[0] + fwd[1] * fwd[1]) + fwd[2] * fwd[2])
    fwd = fwd * f32(f32(1.0) / np.sqrt(d))  # glm::normalize
    pos = np.array(position, dtype=f32)
    look_at = pos + f32(focus_distance) * fwd
    aspect = float(f32(width) / f32(height))
    return create_camera(pos, look_at, aperture, focus_distance, degrees_to_radians(vfov_degrees), aspect)


def camera_to_array(cam: capi.Camera) -> np.ndarray:
    return np.frombuffer(bytes(cam), dtype="<f4").copy()


# sky --------------------------------------------------------------------------------------------------
@dataclass
class Sky:
    turbidity: float = 1.0
    albedo: tuple = (1.0, 1.0, 1.0)
    sun_zenith_degrees: float = 30.0
    sun_azimuth_degrees: float = 0.0

    def to_c(self) -> capi.Sky:
        return capi.Sky(self.turbidity, _f3(self.albedo), self.sun_zenith_degrees, self.sun_azimuth_degrees)


def sky_state(sky: Sky) -> np.ndarray:
    """``AlignedSkyState(sky)`` as 40 floats (params[27], skyRadiances[3], solarRadiances[3], pad[3],
    sunDirection[3], pad)."""
    out = capi.SkyState()
    c_sky = sky.to_c()
    check(lib().rf_sky_state_new(C.byref(c_sky), C.byref(out)))
    return np.frombuffer(bytes(out), dtype="<f4").copy()


# BVH ---------------------------------------------------------------------------------------------------
def build_bvh(triangles: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """``buildBvh``: triangles (n, 3, 3) float32 -> (nodes[BVH_NODE_DTYPE], triangleIndices[u64])."""
    tris = np.ascontiguousarray(triangles, dtype="<f4").reshape(-1, 9)
    n = tris.shape[0]
    nodes = np.zeros(max(2 * n - 1, 1), dtype=BVH_NODE_DTYPE)
    indices = np.zeros(n, dtype=np.uint64)
    num = C.c_uint64(0)
    check(lib().rf_build_bvh(_ptr(tris), n, _ptr(nodes), C.byref(num), _ptr(indices)))
    return nodes[: num.value].copy(), indices


def build_bvh_device(triangles: np.ndarray, device: int = -1) -> tuple[np.ndarray, np.ndarray, float]:
    """``buildBvh`` on the GPU (csrc/bvh_build.cu): byte-identical to :func:`build_bvh`; also returns the device time in ms."""
    tris = np.ascontiguousarray(triangles, dtype="<f4").reshape(-1, 9)
    n = tris.shape[0]
    nodes = np.zeros(max(2 * n - 1, 1), dtype=BVH_NODE_DTYPE)
    indices = np.zeros(n, dtype=np.uint64)
    num = C.c_uint64(0)
    ms = C.c_float(0.0)
    check(lib().rf_build_bvh_device(_ptr(tris), n, device, _ptr(nodes), C.byref(num), _ptr(indices), C.byref(ms)))
    return nodes[: num.value].copy(), indices, ms.value


def reorder_attributes(attributes: np.ndarray, triangle_indices: np.ndarray) -> np.ndarray:
    out = np.empty_like(attributes)
    out[triangle_indices.astype(np.int64)] = attributes
    return out


# .pt container ---------------------------------------------------------------------------------------------
class PtFormat:
    """``nlrs::PtFormat``: the arrays of pt_format.hpp:23-39 as numpy arrays + BGRA8 textures."""

    ARRAYS = {
        "bvh_nodes": (0, BVH_NODE_DTYPE),
        "bvh_position_attributes": (1, POSITIONS_DTYPE),
        "triangle_position_attributes": (2, POSITION_ATTRIBUTE_DTYPE),
        "triangle_vertex_attributes": (3, VERTEX_ATTRIBUTES_DTYPE),
        "vertex_positions": (4, np.dtype(("<f4", 4))),
        "vertex_normals": (5, np.dtype(("<f4", 4))),
        "vertex_tex_coords": (6, np.dtype(("<f4", 2))),
        "vertex_indices": (7, np.dtype("<u4")),
        "model_vertex_positions": (8, np.dtype(("<u8", 2))),
        "model_vertex_normals": (9, np.dtype(("<u8", 2))),
        "model_vertex_tex_coords": (10, np.dtype(("<u8", 2))),
        "model_vertex_indices": (11, np.dtype(("<u8", 2))),
        "model_base_color_texture_indices": (12, np.dtype("<u4")),
    }

    def __init__(self):
        for name, (_, dt) in self.ARRAYS.items():
            setattr(self, name, np.zeros((0,) + dt.shape, dtype=dt.base))
        self.base_color_textures: list[np.ndarray] = []  # each (height, width) uint32 BGRA

    # -- deserialize / serialize through the C-ABI -------------------------------------------------------
    @classmethod
    def _from_handle(cls, handle: C.c_void_p) -> "PtFormat":
        self = cls()
        try:
            for name, (which, dt) in cls.ARRAYS.items():
                data, count, esize = C.c_void_p(), C.c_uint64(), C.c_uint64()
                check(lib().rf_pt_array(handle, which, C.byref(data), C.byref(count), C.byref(esize)))
                assert esize.value == dt.itemsize
                if count.value:
                    buf = C.string_at(data.value, count.value * esize.value)
                    arr = np.frombuffer(buf, dtype=dt.base).reshape((count.value,) + dt.shape).copy()
                else:
                    arr = np.zeros((0,) + dt.shape, dtype=dt.base)
                setattr(self, name, arr)
            for i in range(lib().rf_pt_num_textures(handle)):
                t = capi.Texture()
                check(lib().rf_pt_texture(handle, i, C.byref(t)))
                n = t.width * t.height
                px = np.frombuffer(C.string_at(C.cast(t.pixels, C.c_void_p).value, 4 * n), dtype="<u4") if n else np.zeros(0, "<u4")
                self.base_color_textures.append(px.reshape(t.height, t.width).copy())
        finally:
            lib().rf_pt_destroy(handle)
        return self

    @classmethod
    def load(cls, path) -> "PtFormat":
        handle = C.c_void_p()
        check(lib().rf_pt_load(str(path).encode(), C.byref(handle)))
        return cls._from_handle(handle)

    @classmethod
    def loads(cls, data: bytes) -> "PtFormat":
        handle = C.c_void_p()
        buf = (C.c_char * len(data)).from_buffer_copy(data) if data else (C.c_char * 1)()
        check(lib().rf_pt_load_memory(C.cast(buf, C.c_void_p), len(data), C.byref(handle)))
        return cls._from_handle(handle)

    def _to_handle(self) -> C.c_void_p:
        handle = C.c_void_p()
        check(lib().rf_pt_create(C.byref(handle)))
        try:
            for name, (which, dt) in self.ARRAYS.items():
                arr = np.ascontiguousarray(getattr(self, name), dtype=dt.base)
                count = arr.size * arr.itemsize // dt.itemsize
                check(lib().rf_pt_set_array(handle, which, _ptr(arr) if count else None, count))
            for tex in self.base_color_textures:
                t = np.ascontiguousarray(tex, dtype="<u4")
                check(lib().rf_pt_add_texture(handle, _ptr(t), t.shape[1], t.shape[0]))
        except Exception:
            lib().rf_pt_destroy(handle)
            raise
        return handle

    def save(self, path) -> None:
        handle = self._to_handle()
        try:
            check(lib().rf_pt_save(handle, str(path).encode()))
        finally:
            lib().rf_pt_destroy(handle)

    def dumps(self) -> bytes:
        handle = self._to_handle()
        try:
            size = C.c_uint64()
            check(lib().rf_pt_save_memory(handle, None, 0, C.byref(size)))
            buf = (C.c_char * max(size.value, 1))()
            check(lib().rf_pt_save_memory(handle, C.cast(buf, C.c_void_p), size.value, C.byref(size)))
            return bytes(buf[: size.value])
        finally:
            lib().rf_pt_destroy(handle)


# renderer -----------------------------------------------------------------------------------------------------
@dataclass
class SamplingParams:
    num_samples_per_pixel: int = 128
    num_bounces: int = 4


@dataclass
class RenderParameters:
    framebuffer_size: tuple
    camera: capi.Camera
    sampling_params: SamplingParams = field(default_factory=SamplingParams)
    sky: Sky = field(default_factory=Sky)
    exposure: float = 1.0

    def to_c(self) -> capi.RenderParameters:
        return capi.RenderParameters(
            int(self.framebuffer_size[0]), int(self.framebuffer_size[1]), self.camera,
            capi.SamplingParams(self.sampling_params.num_samples_per_pixel, self.sampling_params.num_bounces),
            self.sky.to_c(), self.exposure)


@dataclass
class SceneArrays:
    """``nlrs::Scene``: the four spans the renderer copies to the device."""
    bvh_nodes: np.ndarray
    position_attributes: np.ndarray
    vertex_attributes: np.ndarray
    base_color_textures: list

    @classmethod
    def from_pt(cls, pt: PtFormat) -> "SceneArrays":  # pt/main.cpp:150-155
        return cls(pt.bvh_nodes, pt.triangle_position_attributes, pt.triangle_vertex_attributes,
                   pt.base_color_textures)


class ReferencePathTracer:
    """Drop-in for ``nlrs::ReferencePathTracer`` running the sm_100a wavefront kernels."""

    def __init__(self, renderer_desc_params: RenderParameters, max_framebuffer_size: tuple, scene: SceneArrays,
                 device: int = -1):
        self._handle = C.c_void_p()
        nodes = np.ascontiguousarray(scene.bvh_nodes)
        pos = np.ascontiguousarray(scene.position_attributes)
        vat = np.ascontiguousarray(scene.vertex_attributes)
        texs = [np.ascontiguousarray(t, dtype="<u4") for t in scene.base_color_textures]
        c_tex = (capi.Texture * max(len(texs), 1))()
        for i, t in enumerate(texs):
            c_tex[i] = capi.Texture(C.cast(_ptr(t), C.POINTER(C.c_uint32)), t.shape[1], t.shape[0])
        c_scene = capi.Scene(_ptr(nodes), nodes.size, _ptr(pos), pos.size, _ptr(vat), vat.size, c_tex, len(texs))
        desc = capi.RendererDescriptor(renderer_desc_params.to_c(), int(max_framebuffer_size[0]), int(max_framebuffer_size[1]))
        check(lib().rf_renderer_create(C.byref(desc), C.byref(c_scene), device, C.byref(self._handle)))
        self._params = renderer_desc_params

    def close(self) -> None:
        if getattr(self, "_handle", None) and self._handle.value:
            lib().rf_renderer_destroy(self._handle)
            self._handle = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_render_parameters(self, params: RenderParameters) -> None:
        c = params.to_c()
        check(lib().rf_renderer_set_render_parameters(self._handle, C.byref(c)))
        self._params = params

    def render(self) -> None:
        check(lib().rf_renderer_render(self._handle))

    def average_renderpass_duration_ms(self) -> float:
        return lib().rf_renderer_average_renderpass_duration_ms(self._handle)

    def render_progress_percentage(self) -> float:
        return lib().rf_renderer_render_progress_percentage(self._handle)

    # -- additions over the reference interface ---------------------------------------------------------
    def read_hdr(self, out: np.ndarray | None = None) -> tuple[np.ndarray, int]:
        w, h = self._params.framebuffer_size
        if out is None:
            out = np.empty((h, w, 4), dtype=np.float32)
        acc = C.c_uint32()
        check(lib().rf_renderer_read_hdr(self._handle, _ptr(out), out.size, C.byref(acc)))
        return out, acc.value

    def read_display(self) -> np.ndarray:
        w, h = self._params.framebuffer_size
        out = np.empty((h, w), dtype=np.uint32)
        check(lib().rf_renderer_read_display(self._handle, _ptr(out), out.size))
        return out

    def render_deferred_lighting(self, inverse_view_projection: np.ndarray, camera_eye, frame_count: int, albedo: np.ndarray,
                                 normal: np.ndarray, depth: np.ndarray, sky: "Sky | None" = None, exposure: float = 1.0) -> None:
        """One frame of the deferred renderer's lighting + resolve passes (csrc/deferred.cuh).  ``inverse_view_projection``:
        (4, 4) with ``m[c]`` = column c (glm / WGSL order); G-buffer: albedo / encoded normal (h, w, 4) float32, reverse-Z
        depth (h, w) float32 (0 = sky)."""
        h, w = depth.shape
        p = capi.DeferredLightingParams()
        p.inverse_view_reverse_z_projection[:] = [float(x) for x in np.asarray(inverse_view_projection, dtype=np.float32).reshape(16)]
        p.camera_eye[:] = [float(camera_eye[0]), float(camera_eye[1]), float(camera_eye[2]), 1.0]
        p.framebuffer_width, p.framebuffer_height, p.frame_count, p.exposure = w, h, frame_count, exposure
        p.sky = (sky or Sky()).to_c()
        a = np.ascontiguousarray(albedo, dtype=np.float32).reshape(h, w, 4)
        n = np.ascontiguousarray(normal, dtype=np.float32).reshape(h, w, 4)
        d = np.ascontiguousarray(depth, dtype=np.float32)
        check(lib().rf_renderer_render_deferred_lighting(self._handle, C.byref(p), _ptr(a), _ptr(n), _ptr(d)))
        self._deferred_size = (w, h)

    def read_deferred(self):
        """(sampleBuffer (h, w, 3), accumulationBuffer (h, w, 3), BGRA8 display (h, w)) of the last deferred frame."""
        w, h = self._deferred_size
        sample = np.empty((h, w, 3), dtype=np.float32)
        accumulation = np.empty((h, w, 3), dtype=np.float32)
        display = np.empty((h, w), dtype=np.uint32)
        check(lib().rf_renderer_read_deferred(self._handle, _ptr(sample), _ptr(accumulation), _ptr(display)))
        return sample, accumulation, display

    def hdr_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib().rf_renderer_hdr_ipc_handle(self._handle, buf))
        return buf.raw

    def set_hdr_peer(self, handle: "bytes | None") -> None:
        check(lib().rf_renderer_set_hdr_peer(self._handle, C.create_string_buffer(handle, 64) if handle is not None else None))

    def hdr_device_ptr(self) -> int:
        return lib().rf_renderer_hdr_device_ptr(self._handle) or 0

    def exchange_device_ptr(self) -> int:
        """Root of a peer-memory exchange: device pointer of the last complete exchanged frame (0 otherwise)."""
        return lib().rf_renderer_exchange_device_ptr(self._handle) or 0

    def set_option(self, name: str, value: int) -> None:
        check(lib().rf_renderer_set_option(self._handle, name.encode(), int(value)))

    def set_stream(self, cuda_stream: int) -> None:
        check(lib().rf_renderer_set_stream(self._handle, C.c_void_p(cuda_stream)))

    def synchronize(self) -> None:
        check(lib().rf_renderer_synchronize(self._handle))

    def set_frame_count(self, n: int) -> None:
        check(lib().rf_renderer_set_frame_count(self._handle, n))

    @property
    def frame_count(self) -> int:
        return lib().rf_renderer_frame_count(self._handle)

    @property
    def accumulated_sample_count(self) -> int:
        return lib().rf_renderer_accumulated_sample_count(self._handle)

    def set_tile_partition(self, rank: int, world: int) -> None:
        check(lib().rf_renderer_set_tile_partition(self._handle, rank, world))

    def stats(self) -> dict:
        s = capi.FrameStats()
        check(lib().rf_renderer_get_stats(self._handle, C.byref(s)))
        return {name: getattr(s, name) for name, _ in capi.FrameStats._fields_}

    def reset_stats(self) -> None:
        check(lib().rf_renderer_reset_stats(self._handle))

    def set_stage_timing(self, enabled: bool) -> None:
        check(lib().rf_renderer_set_stage_timing(self._handle, int(enabled)))

    def set_tuning(self, tri_min: int = 0, refill_min: int = 0, blocks_per_sm: int = 0) -> None:
        check(lib().rf_renderer_set_tuning(self._handle, tri_min, refill_min, blocks_per_sm))

    def set_pipeline(self, sub_frames: int = 0, persistent_kernel: int = -1, variant: int = -1, block_threads: int = 0) -> None:
        check(lib().rf_renderer_set_pipeline(self._handle, sub_frames, persistent_kernel, variant, block_threads))

    def set_tail_policy(self, evict_max: int = -1) -> None:
        check(lib().rf_renderer_set_tail_policy(self._handle, evict_max))


class TraversalScene:
    """Device-resident (bvhNodes, triangles) for the GPU twin of ``rayIntersectBvh``."""

    def __init__(self, bvh_nodes: np.ndarray, triangles: np.ndarray, device: int = -1):
        self._handle = C.c_void_p()
        nodes = np.ascontiguousarray(bvh_nodes)
        tris = np.ascontiguousarray(triangles, dtype="<f4").reshape(-1, 9) if triangles.dtype != POSITIONS_DTYPE else np.ascontiguousarray(triangles)
        n_tris = tris.shape[0]
        check(lib().rf_traversal_scene_create(_ptr(nodes), nodes.size, _ptr(tris), n_tris, device, C.byref(self._handle)))

    def close(self) -> None:
        if getattr(self, "_handle", None) and self._handle.value:
            lib().rf_traversal_scene_destroy(self._handle)
            self._handle = C.c_void_p()

    __del__ = close

    def set_kernel(self, trace_kernel: int) -> None:
        """0 / 1: one node per visit (default); 2: child-pair records."""
        check(lib().rf_traversal_scene_set_kernel(self._handle, trace_kernel))

    def ray_intersect_bvh(self, rays: np.ndarray, ray_t_max: float):
        """rays (n, 6) -> (hit bool[n], p_t float32[n, 4], nodes_visited uint32[n])."""
        rays = np.ascontiguousarray(rays, dtype="<f4").reshape(-1, 6)
        n = rays.shape[0]
        hit = np.zeros(n, dtype=np.uint8)
        p_t = np.zeros((n, 4), dtype=np.float32)
        nodes = np.zeros(n, dtype=np.uint32)
        check(lib().rf_ray_intersect_bvh(self._handle, _ptr(rays), n, ray_t_max, _ptr(hit), _ptr(p_t), _ptr(nodes)))
        return hit.astype(bool), p_t, nodes

    def pick_focus_distance(self, camera: capi.Camera, position, forward, cursor_x: float, cursor_y: float, window_width: int,
                            window_height: int):
        """Click-to-focus (pt/main.cpp:198-226): the new focus distance, or None for a miss / a cursor outside the window."""
        pos = (C.c_float * 3)(*[float(x) for x in position])
        fwd = (C.c_float * 3)(*[float(x) for x in forward])
        hit, dist = C.c_uint8(0), C.c_float(0.0)
        check(lib().rf_pick_focus_distance(self._handle, C.byref(camera), C.byref(pos), C.byref(fwd), cursor_x, cursor_y, window_width,
                                           window_height, C.byref(hit), C.byref(dist)))
        return dist.value if hit.value else None

    def bvh_visualizer_node_counts(self, camera: capi.Camera, width: int, height: int, ray_t_max: float = FLT_MAX):
        out = np.zeros((height, width), dtype=np.uint32)
        ms = C.c_float()
        check(lib().rf_bvh_visualizer_node_counts(self._handle, C.byref(camera), width, height, ray_t_max, _ptr(out), C.byref(ms)))
        return out, ms.value


def bvh_visualizer_camera(bvh_nodes: np.ndarray, width: int, height: int, float_constants: bool = False) -> capi.Camera:
    """Camera of bvh-visualizer/main.cpp:36-55 (root AABB, eye = centroid - (-0.8 d, 0, 0.8 d), vfov 70 degrees, focus 1).
    The x offset is the *double* product ``-0.8 * d`` rounded to float (main.cpp:49); tests/bvh.cpp:67 uses
    ``-0.8f`` instead (``float_constants=True``)."""
    f32 = np.float32
    lo = bvh_nodes["aabb_min"][0].astype(f32)
    hi = bvh_nodes["aabb_max"][0].astype(f32)
    diag = hi - lo
    centroid = f32(0.5) * (lo + hi)
    if diag[0] > diag[1] and diag[0] > diag[2]:
        d = diag[0]
    elif diag[1] > diag[2]:
        d = diag[1]
    else:
        d = diag[2]
    ox = f32(f32(-0.8) * d) if float_constants else f32(-0.8 * float(d))
    oz = f32(f32(0.8) * d)
    eye = centroid - np.array([ox, f32(0.0), oz], dtype=f32)
    aspect = float(f32(width) / f32(height))
    return create_camera(eye, centroid, 0.0, 1.0, degrees_to_radians(70.0), aspect)


__all__ = [
    "BVH_NODE_DTYPE", "POSITIONS_DTYPE", "POSITION_ATTRIBUTE_DTYPE", "VERTEX_ATTRIBUTES_DTYPE", "FLT_MAX",
    "RayfinderError", "Sky", "SamplingParams", "RenderParameters", "SceneArrays", "PtFormat",
    "ReferencePathTracer", "TraversalScene", "create_camera", "fly_camera", "camera_to_array",
    "degrees_to_radians", "sky_state", "build_bvh", "build_bvh_device", "reorder_attributes", "bvh_visualizer_camera",
]

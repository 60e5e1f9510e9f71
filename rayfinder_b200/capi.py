"""ctypes binding of include/rayfinder_b200.h — the stub a Python host would use.

The shared library is built in-tree by ``rayfinder_b200._build`` (nvcc, sm_100a).  There is no
Python/numpy implementation of any entry point: if the library is missing this module raises, and
on a machine without a CUDA device the renderer constructors raise ``RayfinderError``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

_PKG = Path(__file__).resolve().parent
# RAYFINDER_B200_LIB selects another build of the same library (kernel experiments); there is still no fallback.
LIB_PATH = Path(os.environ.get("RAYFINDER_B200_LIB", _PKG / "librayfinder_b200.so"))

RF_OK = 0
RF_ERROR_INVALID_ARGUMENT = 1
RF_ERROR_CUDA = 2
RF_ERROR_IO = 3
RF_ERROR_FORMAT = 4
RF_ERROR_OUT_OF_RANGE = 5


class RayfinderError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


class BvhNode(C.Structure):  # common/bvh.hpp:13-21
    _fields_ = [("aabb_min", C.c_float * 3), ("pad0", C.c_float), ("aabb_max", C.c_float * 3), ("pad1", C.c_float),
                ("triangles_offset", C.c_uint32), ("second_child_offset", C.c_uint32),
                ("triangle_count", C.c_uint32), ("split_axis", C.c_uint32)]


class Texture(C.Structure):  # common/texture.hpp
    _fields_ = [("pixels", C.POINTER(C.c_uint32)), ("width", C.c_uint32), ("height", C.c_uint32)]


class Scene(C.Structure):  # pt/reference_path_tracer.hpp:45-51
    _fields_ = [("bvh_nodes", C.c_void_p), ("num_bvh_nodes", C.c_uint64),
                ("position_attributes", C.c_void_p), ("num_position_attributes", C.c_uint64),
                ("vertex_attributes", C.c_void_p), ("num_vertex_attributes", C.c_uint64),
                ("base_color_textures", C.POINTER(Texture)), ("num_base_color_textures", C.c_uint64)]


class Camera(C.Structure):  # common/camera.hpp:10-21
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("up", C.c_float * 3), ("right", C.c_float * 3),
                ("lens_radius", C.c_float)]


class SamplingParams(C.Structure):  # pt/reference_path_tracer.hpp:26-32
    _fields_ = [("num_samples_per_pixel", C.c_uint32), ("num_bounces", C.c_uint32)]


class Sky(C.Structure):  # pt/aligned_sky_state.hpp:15-23
    _fields_ = [("turbidity", C.c_float), ("albedo", C.c_float * 3), ("sun_zenith_degrees", C.c_float),
                ("sun_azimuth_degrees", C.c_float)]


class RenderParameters(C.Structure):  # pt/reference_path_tracer.hpp:34-43
    _fields_ = [("framebuffer_width", C.c_uint32), ("framebuffer_height", C.c_uint32), ("camera", Camera),
                ("sampling_params", SamplingParams), ("sky", Sky), ("exposure", C.c_float)]


class DeferredLightingParams(C.Structure):  # deferred_renderer_lighting_pass.wgsl:8-13 + resolve pass Uniforms
    _fields_ = [("inverse_view_reverse_z_projection", C.c_float * 16), ("camera_eye", C.c_float * 4),
                ("framebuffer_width", C.c_uint32), ("framebuffer_height", C.c_uint32), ("frame_count", C.c_uint32),
                ("exposure", C.c_float), ("sky", Sky)]


class RendererDescriptor(C.Structure):  # pt/reference_path_tracer.hpp:53-57
    _fields_ = [("render_params", RenderParameters), ("max_framebuffer_width", C.c_int32),
                ("max_framebuffer_height", C.c_int32)]


class SkyState(C.Structure):  # pt/aligned_sky_state.hpp:34-41
    _fields_ = [("params", C.c_float * 27), ("sky_radiances", C.c_float * 3), ("solar_radiances", C.c_float * 3),
                ("padding1", C.c_float * 3), ("sun_direction", C.c_float * 3), ("padding2", C.c_float)]


class FrameStats(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("paths", C.c_uint64), ("closest_rays", C.c_uint64),
                ("shadow_rays", C.c_uint64), ("closest_nodes_visited", C.c_uint64),
                ("closest_triangles_tested", C.c_uint64), ("shadow_nodes_visited", C.c_uint64),
                ("shadow_triangles_tested", C.c_uint64), ("device_ms_total", C.c_double),
                ("device_ms_trace", C.c_double), ("device_ms_shade", C.c_double), ("device_ms_other", C.c_double),
                ("kernel_launches", C.c_uint64), ("sub_frames", C.c_uint32), ("evict_max", C.c_uint32),
                ("node_records_loaded", C.c_uint64), ("trace_kernel", C.c_uint32), ("persistent_kernel", C.c_uint32)]


# name -> (restype, argtypes); every symbol include/rayfinder_b200.h declares.
_P = C.c_void_p
SIGNATURES = {
    "rf_last_error": (C.c_char_p, []),
    "rf_renderer_create": (C.c_int32, [C.POINTER(RendererDescriptor), C.POINTER(Scene), C.c_int32, C.POINTER(_P)]),
    "rf_renderer_destroy": (None, [_P]),
    "rf_renderer_set_render_parameters": (C.c_int32, [_P, C.POINTER(RenderParameters)]),
    "rf_renderer_render": (C.c_int32, [_P]),
    "rf_renderer_average_renderpass_duration_ms": (C.c_float, [_P]),
    "rf_renderer_render_progress_percentage": (C.c_float, [_P]),
    "rf_renderer_read_hdr": (C.c_int32, [_P, _P, C.c_uint64, C.POINTER(C.c_uint32)]),
    "rf_renderer_read_display": (C.c_int32, [_P, _P, C.c_uint64]),
    "rf_renderer_hdr_device_ptr": (_P, [_P]),
    "rf_renderer_set_stream": (C.c_int32, [_P, _P]),
    "rf_renderer_synchronize": (C.c_int32, [_P]),
    "rf_renderer_set_frame_count": (C.c_int32, [_P, C.c_uint32]),
    "rf_renderer_frame_count": (C.c_uint32, [_P]),
    "rf_renderer_accumulated_sample_count": (C.c_uint32, [_P]),
    "rf_renderer_set_tile_partition": (C.c_int32, [_P, C.c_uint32, C.c_uint32]),
    "rf_renderer_get_stats": (C.c_int32, [_P, C.POINTER(FrameStats)]),
    "rf_renderer_reset_stats": (C.c_int32, [_P]),
    "rf_renderer_set_stage_timing": (C.c_int32, [_P, C.c_int32]),
    "rf_renderer_set_tuning": (C.c_int32, [_P, C.c_uint32, C.c_uint32, C.c_uint32]),
    "rf_renderer_set_pipeline": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "rf_renderer_set_tail_policy": (C.c_int32, [_P, C.c_int32]),
    "rf_renderer_hdr_ipc_handle": (C.c_int32, [_P, _P]),
    "rf_renderer_set_hdr_peer": (C.c_int32, [_P, _P]),
    "rf_renderer_exchange_device_ptr": (_P, [_P]),
    "rf_renderer_set_option": (C.c_int32, [_P, C.c_char_p, C.c_int64]),
    "rf_renderer_render_deferred_lighting": (C.c_int32, [_P, C.POINTER(DeferredLightingParams), _P, _P, _P]),
    "rf_renderer_read_deferred": (C.c_int32, [_P, _P, _P, _P]),
    "rf_traversal_scene_create": (C.c_int32, [_P, C.c_uint64, _P, C.c_uint64, C.c_int32, C.POINTER(_P)]),
    "rf_traversal_scene_destroy": (None, [_P]),
    "rf_traversal_scene_set_kernel": (C.c_int32, [_P, C.c_int32]),
    "rf_ray_intersect_bvh": (C.c_int32, [_P, _P, C.c_uint64, C.c_float, _P, _P, _P]),
    "rf_bvh_visualizer_node_counts": (C.c_int32, [_P, C.POINTER(Camera), C.c_uint32, C.c_uint32, C.c_float, _P, C.POINTER(C.c_float)]),
    "rf_pick_focus_distance": (C.c_int32, [_P, C.POINTER(Camera), C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_double, C.c_double,
                               C.c_int32, C.c_int32, C.POINTER(C.c_uint8), C.POINTER(C.c_float)]),
    "rf_create_camera": (C.c_int32, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(Camera)]),
    "rf_sky_state_new": (C.c_int32, [C.POINTER(Sky), C.POINTER(SkyState)]),
    "rf_build_bvh": (C.c_int32, [_P, C.c_uint64, _P, C.POINTER(C.c_uint64), _P]),
    "rf_build_bvh_device": (C.c_int32, [_P, C.c_uint64, C.c_int32, _P, C.POINTER(C.c_uint64), _P, C.POINTER(C.c_float)]),
    "rf_build_bvh_device_set_mode": (None, [C.c_int32]),
    "rf_build_bvh_device_last_phases": (C.c_uint32, [C.POINTER(C.c_float)]),
    "rf_build_bvh_device_release": (None, []),
    "rf_pt_create": (C.c_int32, [C.POINTER(_P)]),
    "rf_pt_destroy": (None, [_P]),
    "rf_pt_load": (C.c_int32, [C.c_char_p, C.POINTER(_P)]),
    "rf_pt_load_memory": (C.c_int32, [_P, C.c_uint64, C.POINTER(_P)]),
    "rf_pt_save": (C.c_int32, [_P, C.c_char_p]),
    "rf_pt_save_memory": (C.c_int32, [_P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "rf_pt_array": (C.c_int32, [_P, C.c_int32, C.POINTER(_P), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "rf_pt_set_array": (C.c_int32, [_P, C.c_int32, _P, C.c_uint64]),
    "rf_pt_num_textures": (C.c_uint64, [_P]),
    "rf_pt_texture": (C.c_int32, [_P, C.c_uint64, C.POINTER(Texture)]),
    "rf_pt_add_texture": (C.c_int32, [_P, _P, C.c_uint32, C.c_uint32]),
    "rf_pt_scene": (C.c_int32, [_P, C.POINTER(Scene), C.POINTER(Texture)]),
    "rf_texture_from_memory": (C.c_int32, [_P, C.c_uint64, _P, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "rf_bake_gltf": (C.c_int32, [C.c_char_p, C.POINTER(_P)]),
    "rf_has_cuda_kernels": (C.c_int32, []),
    "rf_build_info": (C.c_char_p, []),
}

_lib = None


def lib() -> C.CDLL:
    """Load librayfinder_b200.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m rayfinder_b200._build` "
                "(nvcc, sm_100a).  rayfinder_b200 has no Python or CPU fallback.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != RF_OK:
        raise RayfinderError(status, lib().rf_last_error().decode("utf-8", "replace"))

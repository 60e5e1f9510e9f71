"""ctypes access to the parity oracle (oracle/liboracle.so) and to the compiled reference CPU path
(oracle/_ref/libref_oracle.so).  TEST INFRASTRUCTURE: imported only by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs — never by rayfinder_b200/."""
from __future__ import annotations

import ctypes as C
import lzma
import os
import subprocess
from functools import lru_cache
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liboracle.so"
REF_LIB = ORACLE_DIR / "_ref" / "libref_oracle.so"
GOLDEN = ROOT / "tests" / "golden"
ASSETS = ROOT / "assets"
DATA = ROOT / "rayfinder_b200" / "data"

_P = C.c_void_p


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def build_oracle() -> None:
    if not ORACLE_LIB.exists() or ORACLE_LIB.stat().st_mtime < (ORACLE_DIR / "oracle.cpp").stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)


class OracleScene(C.Structure):
    _fields_ = [("nodes", _P), ("positionAttributes", _P), ("vertexAttributes", _P), ("texDesc", _P),
                ("numTextures", C.c_uint32), ("texels", _P), ("numTexels", C.c_uint64), ("blueNoiseRg8", _P)]


class OracleFrame(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("frameCount", C.c_uint32),
                ("numSamplesPerPixel", C.c_uint32), ("numBounces", C.c_uint32),
                ("accumulatedSampleCount", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32),
                ("camera", C.c_float * 19), ("skyState", C.c_float * 40)]


class OracleDeferred(C.Structure):
    _fields_ = [("inverseViewReverseZProjection", C.c_float * 16), ("cameraEye", C.c_float * 4), ("width", C.c_uint32),
                ("height", C.c_uint32), ("frameCount", C.c_uint32), ("skyState", C.c_float * 40)]


@lru_cache(maxsize=None)
def oracle() -> C.CDLL:
    build_oracle()
    lib = C.CDLL(str(ORACLE_LIB))
    lib.oracle_bvh_visualizer.restype = C.c_double
    lib.oracle_bvh_visualizer.argtypes = [_P, _P, _P, C.c_int, C.c_int, C.c_float, _P, C.c_int]
    lib.oracle_bvh_visualizer_rows.restype = C.c_double
    lib.oracle_bvh_visualizer_rows.argtypes = [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, C.c_int]
    lib.oracle_intersect_batch.argtypes = [_P, _P, _P, C.c_uint64, C.c_float, _P, _P, _P, C.c_int]
    lib.oracle_create_camera.argtypes = [_P, _P, C.c_float, C.c_float, C.c_float, C.c_float, _P]
    lib.oracle_ray_intersect_aabb.argtypes = [_P, _P, C.c_float]
    lib.oracle_ray_intersect_triangle.argtypes = [_P, _P, C.c_float, _P]
    lib.oracle_sky_radiance.restype = C.c_float
    lib.oracle_sky_radiance.argtypes = [_P, C.c_float, C.c_float, C.c_uint32]
    lib.oracle_solar_constants.argtypes = [_P]
    lib.oracle_render_frame.restype = C.c_double
    lib.oracle_render_frame.argtypes = [C.POINTER(OracleScene), C.POINTER(OracleFrame), _P, _P, _P, C.c_int]
    lib.oracle_frame_rays.restype = C.c_uint64
    lib.oracle_frame_rays.argtypes = [C.POINTER(OracleScene), C.POINTER(OracleFrame), C.c_uint32, _P, _P, C.c_uint64, C.c_int]
    lib.oracle_display.argtypes = [_P, C.c_uint64, C.c_float, C.c_float, _P]
    lib.oracle_deferred_lighting.restype = C.c_double
    lib.oracle_deferred_lighting.argtypes = [C.POINTER(OracleScene), C.POINTER(OracleDeferred), _P, _P, _P, _P, _P, C.c_int]
    lib.oracle_deferred_resolve.argtypes = [_P, _P, C.c_uint64, C.c_uint32]
    return lib


def have_ref() -> bool:
    return REF_LIB.exists()


@lru_cache(maxsize=None)
def ref() -> C.CDLL:
    lib = C.CDLL(str(REF_LIB))
    lib.ref_bvh_build.restype = _P
    lib.ref_bvh_build.argtypes = [_P, C.c_uint64]
    lib.ref_bvh_num_nodes.restype = C.c_uint64
    lib.ref_bvh_num_nodes.argtypes = [_P]
    lib.ref_bvh_copy.argtypes = [_P, _P, _P]
    lib.ref_bvh_free.argtypes = [_P]
    lib.ref_create_camera.argtypes = [_P, _P, C.c_float, C.c_float, C.c_float, C.c_float, _P]
    lib.ref_generate_camera_ray.argtypes = [_P, C.c_float, C.c_float, _P]
    lib.ref_ray_intersect_aabb.argtypes = [_P, _P, C.c_float]
    lib.ref_ray_intersect_triangle.argtypes = [_P, _P, C.c_float, _P]
    lib.ref_intersect_batch.argtypes = [_P, C.c_uint64, _P, C.c_uint64, _P, C.c_uint64, C.c_float, _P, _P, _P, C.c_int]
    lib.ref_bvh_visualizer.restype = C.c_double
    lib.ref_bvh_visualizer.argtypes = [_P, C.c_uint64, _P, C.c_uint64, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, C.c_int]
    lib.ref_sky_state_new.argtypes = [C.c_float, C.c_float, _P, _P]
    lib.ref_sky_state_radiance.restype = C.c_float
    lib.ref_sky_state_radiance.argtypes = [_P, C.c_float, C.c_float, C.c_int]
    lib.ref_blue_noise.restype = C.POINTER(C.c_uint8)
    lib.ref_blue_noise.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    return lib


def num_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def blue_noise_rg8() -> np.ndarray:
    return np.fromfile(DATA / "blue_noise_128x128_rg8.bin", dtype=np.uint8)


# ---- scenes ------------------------------------------------------------------------------------------
def duck_pt_bytes() -> bytes:
    """Duck.pt baked from the reference's assets/Duck.glb by rayfinder_b200.baker (committed as a fixture,
    see tests/golden/make_golden.py)."""
    return lzma.decompress((GOLDEN / "Duck.pt.xz").read_bytes())


def triangles9(pt) -> np.ndarray:
    """nlrs::Positions view (n, 9) of PtFormat.bvh_position_attributes."""
    return np.ascontiguousarray(pt.bvh_position_attributes).view("<f4").reshape(-1, 9)


def tex_desc(textures) -> tuple[np.ndarray, np.ndarray]:
    """(descriptors (n,3) u32, concatenated texels) as packed by reference_path_tracer.cpp:209-270."""
    desc, off = [], 0
    for t in textures:
        desc.append((t.shape[1], t.shape[0], off))
        off += t.size
    return np.array(desc, dtype=np.uint32), np.concatenate([np.ascontiguousarray(t).reshape(-1) for t in textures]).astype("<u4")


# ---- the reference arm's own scene loading and camera (no product code involved) ----------------------
BVH_NODE_DTYPE = np.dtype([("aabb_min", "<f4", 3), ("pad0", "<f4"), ("aabb_max", "<f4", 3), ("pad1", "<f4"),
                           ("triangles_offset", "<u4"), ("second_child_offset", "<u4"), ("triangle_count", "<u4"),
                           ("split_axis", "<u4")])  # common/bvh.hpp:13-21


class NumpyPt:
    """A .pt file read with numpy alone, following the layout of pt-format/pt_format.cpp:240-269 — the loader of
    bench.py's reference arm, which must not map the product library."""

    ELEM_SIZE = (48, 36, 48, 80, 16, 16, 8, 4, 16, 16, 16, 16, 4)

    def __init__(self, raw: bytes):
        if raw[:9] != b"PTFORMAT3":
            raise ValueError("Invalid file format: expected PtFormat file.")
        pos, arrays = 9, []
        for size in self.ELEM_SIZE:
            n = int(np.frombuffer(raw, "<u8", 1, pos)[0])
            arrays.append(np.frombuffer(raw, np.uint8, n * size, pos + 8))
            pos += 8 + n * size
        self.bvh_nodes = arrays[0].view(BVH_NODE_DTYPE).copy()
        self.bvh_position_attributes = arrays[1].view("<f4").reshape(-1, 9).copy()
        self.triangle_position_attributes = arrays[2].view("<f4").reshape(-1, 12).copy()
        self.triangle_vertex_attributes = arrays[3].view("<f4").reshape(-1, 20).copy()
        self.base_color_textures = []
        num_textures = int(np.frombuffer(raw, "<u8", 1, pos)[0])
        pos += 8
        for _ in range(num_textures):
            w, h = (int(x) for x in np.frombuffer(raw, "<u4", 2, pos))
            n = int(np.frombuffer(raw, "<u8", 1, pos + 8)[0])
            self.base_color_textures.append(np.frombuffer(raw, "<u4", n, pos + 16).reshape(h, w).copy())
            pos += 16 + 4 * n

    @classmethod
    def load_scene(cls, name: str) -> "NumpyPt | None":
        for path in (ASSETS / f"{name}.pt", ASSETS / f"{name}.pt.xz"):
            if path.exists():
                raw = path.read_bytes()
                return cls(lzma.decompress(raw) if path.suffix == ".xz" else raw)
        return None


def _degrees_to_radians(deg: float) -> np.float32:
    return np.float32(np.float32(np.float32(deg) * np.float32(np.pi)) / np.float32(180.0))  # Angle::degrees, units/angle.hpp:12-15


def fly_camera_array(width: int, height: int, position=(1.22, 1.25, -1.25), yaw_degrees=129.64, pitch_degrees=-13.73,
                     vfov_degrees=70.0, aperture=0.0, focus_distance=10.0) -> np.ndarray:
    """The benchmark view (fly_camera_controller.hpp:47-52, cameraOrientation :138-148, vfov 70 deg pt/main.cpp:49,314) as
    the 19 floats of nlrs::Camera, built by the reference's own createCamera (oracle/_ref) or, without it, the port."""
    import math

    f32 = np.float32
    yaw, pitch = _degrees_to_radians(yaw_degrees), _degrees_to_radians(pitch_degrees)
    cy, sy, cp, sp = (f32(fn(float(a))) for fn, a in ((math.cos, yaw), (math.sin, yaw), (math.cos, pitch), (math.sin, pitch)))
    fwd = np.array([cy * cp, sp, sy * cp], dtype=f32)
    d = f32(f32(fwd[0] * fwd[0] + fwd[1] * fwd[1]) + fwd[2] * fwd[2])
    fwd = fwd * f32(f32(1.0) / np.sqrt(d))
    pos = np.array(position, dtype=f32)
    look_at = (pos + f32(focus_distance) * fwd).astype(f32)
    out = np.zeros(19, dtype=f32)
    aspect = float(f32(width) / f32(height))
    if have_ref():  # takes degrees: it calls Angle::degrees itself
        ref().ref_create_camera(_ptr(pos), _ptr(look_at), float(aperture), float(focus_distance), float(vfov_degrees), aspect, _ptr(out))
    else:
        oracle().oracle_create_camera(_ptr(pos), _ptr(look_at), float(aperture), float(focus_distance),
                                      float(_degrees_to_radians(vfov_degrees)), aspect, _ptr(out))
    return out


def default_sky_state() -> np.ndarray:
    """AlignedSkyState(Sky{}) (aligned_sky_state.hpp:17-20, 43-70) as 40 floats, from the reference's sky_state_new when
    oracle/_ref is built."""
    import math

    f32 = np.float32
    zenith, azimuth = _degrees_to_radians(30.0), _degrees_to_radians(0.0)
    out = np.zeros(40, dtype=f32)
    if have_ref():
        state = np.zeros(33, dtype=f32)  # params[27], sky_radiances[3], solar_radiances[3]
        albedo = np.ones(3, dtype=f32)
        ref().ref_sky_state_new(float(f32(0.5) * f32(np.pi) - zenith), 1.0, _ptr(albedo), _ptr(state))
        out[:33] = state
    else:
        raise RuntimeError("default_sky_state needs oracle/_ref")
    sin_z, cos_z = f32(math.sin(float(zenith))), f32(math.cos(float(zenith)))
    d = np.array([sin_z * f32(math.cos(float(azimuth))), cos_z, -sin_z * f32(math.sin(float(azimuth)))], dtype=f32)
    n = f32(f32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    out[36:39] = d * f32(f32(1.0) / np.sqrt(n))
    return out


def frame_rays(pt, width, height, cam19, sky40, spp, bounces, tile_stride=1, frame_count=0, threads=None):
    """(rays (n, 6) f32, kind (n,) u8: 0 closest-hit / 1 shadow) — every ray the path tracer traces for the pixels of every
    ``tile_stride``-th 32x32 tile of the frame (oracle_frame_rays)."""
    o = OracleRenderer(pt, width, height, cam19, sky40, spp, bounces, threads=threads)
    fr = OracleFrame(width, height, frame_count, spp, bounces, 0, 0, 1, (C.c_float * 19)(*o.cam19), (C.c_float * 40)(*o.sky40))
    n = oracle().oracle_frame_rays(C.byref(o.scene), C.byref(fr), tile_stride, None, None, 0, o.threads)
    rays = np.zeros((n, 6), dtype=np.float32)
    kind = np.zeros(n, dtype=np.uint8)
    oracle().oracle_frame_rays(C.byref(o.scene), C.byref(fr), tile_stride, _ptr(rays), _ptr(kind), n, o.threads)
    return rays, kind


# ---- oracle wrappers ---------------------------------------------------------------------------------
def oracle_node_counts(nodes, tris9, cam19, width, height, t_max, threads=None, rows=None):
    out = np.zeros((height, width), dtype=np.uint32)
    cam19 = np.ascontiguousarray(cam19, dtype=np.float32)
    r0, r1 = rows if rows else (0, height)
    secs = oracle().oracle_bvh_visualizer_rows(_ptr(nodes), _ptr(tris9), _ptr(cam19), width, height, r0, r1, t_max, _ptr(out),
                                               threads or num_threads())
    return out, secs


def ref_node_counts(nodes, tris9, cam19, width, height, t_max, threads=None, rows=None):
    out = np.zeros((height, width), dtype=np.uint32)
    cam19 = np.ascontiguousarray(cam19, dtype=np.float32)
    r0, r1 = rows if rows else (0, height)
    secs = ref().ref_bvh_visualizer(_ptr(nodes), nodes.size, _ptr(tris9), tris9.shape[0], _ptr(cam19), width, height,
                                    r0, r1, t_max, _ptr(out), threads or num_threads())
    return out, secs


def _batch(fn_call, n):
    hit = np.zeros(n, dtype=np.uint8)
    p_t = np.zeros((n, 4), dtype=np.float32)
    visited = np.zeros(n, dtype=np.uint32)
    fn_call(hit, p_t, visited)
    return hit.astype(bool), p_t, visited


def oracle_intersect(nodes, tris9, rays, t_max, threads=None):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
    return _batch(lambda h, p, v: oracle().oracle_intersect_batch(
        _ptr(nodes), _ptr(tris9), _ptr(rays), rays.shape[0], t_max, _ptr(h), _ptr(p), _ptr(v), threads or num_threads()), rays.shape[0])


def ref_intersect(nodes, tris9, rays, t_max, threads=None):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
    return _batch(lambda h, p, v: ref().ref_intersect_batch(
        _ptr(nodes), nodes.size, _ptr(tris9), tris9.shape[0], _ptr(rays), rays.shape[0], t_max, _ptr(h), _ptr(p), _ptr(v),
        threads or num_threads()), rays.shape[0])


COUNTER_NAMES = ("paths", "closest_rays", "shadow_rays", "closest_nodes_visited", "closest_triangles_tested",
                 "shadow_nodes_visited", "shadow_triangles_tested")


class OracleRenderer:
    """Oracle B driver with the accumulation bookkeeping of reference_path_tracer.cpp:556-591."""

    def __init__(self, pt, width, height, cam19, sky40, spp, bounces, rank=0, world=1, threads=None):
        self.nodes = np.ascontiguousarray(pt.bvh_nodes)
        self.pos = np.ascontiguousarray(pt.triangle_position_attributes)
        self.vat = np.ascontiguousarray(pt.triangle_vertex_attributes)
        self.desc, self.texels = tex_desc(pt.base_color_textures)
        self.bn = blue_noise_rg8()
        self.scene = OracleScene(_ptr(self.nodes), _ptr(self.pos), _ptr(self.vat), _ptr(self.desc), len(self.desc),
                                 _ptr(self.texels), self.texels.size, _ptr(self.bn))
        self.width, self.height, self.spp, self.bounces = width, height, spp, bounces
        self.cam19 = np.ascontiguousarray(cam19, dtype=np.float32)
        self.sky40 = np.ascontiguousarray(sky40, dtype=np.float32)
        self.rank, self.world = rank, world
        self.threads = threads or num_threads()
        self.frame_count = 0
        self.accumulated = 0
        self.image = np.zeros((height, width, 4), dtype=np.float32)
        self.counters = np.zeros(9, dtype=np.uint64)
        self.path_lengths = np.zeros((height, width), dtype=np.uint8)
        self.seconds = 0.0

    def render(self):
        fr = OracleFrame(self.width, self.height, self.frame_count, self.spp, self.bounces, self.accumulated,
                         self.rank, self.world, (C.c_float * 19)(*self.cam19), (C.c_float * 40)(*self.sky40))
        self.frame_count += 1
        self.seconds += oracle().oracle_render_frame(C.byref(self.scene), C.byref(fr), _ptr(self.image), _ptr(self.counters),
                                                     _ptr(self.path_lengths), self.threads)
        self.accumulated = min(self.accumulated + 1, self.spp)

    def stats(self) -> dict:
        return {k: int(v) for k, v in zip(COUNTER_NAMES, self.counters)}

    def display(self, exposure: float) -> np.ndarray:
        out = np.zeros((self.height, self.width), dtype=np.uint32)
        oracle().oracle_display(_ptr(self.image), self.width * self.height, float(max(self.accumulated, 1)), exposure, _ptr(out))
        return out


class OracleDeferredLighting:
    """The deferred renderer's lighting + resolve passes restated (oracle.cpp, oracle_deferred_lighting / _resolve)."""

    def __init__(self, pt, sky40, threads=None):
        self.nodes = np.ascontiguousarray(pt.bvh_nodes)
        self.pos = np.ascontiguousarray(pt.triangle_position_attributes)
        self.vat = np.ascontiguousarray(pt.triangle_vertex_attributes)
        self.desc, self.texels = tex_desc(pt.base_color_textures)
        self.bn = blue_noise_rg8()
        self.scene = OracleScene(_ptr(self.nodes), _ptr(self.pos), _ptr(self.vat), _ptr(self.desc), len(self.desc),
                                 _ptr(self.texels), self.texels.size, _ptr(self.bn))
        self.sky40 = np.ascontiguousarray(sky40, dtype=np.float32)
        self.threads = threads or num_threads()
        self.counters = np.zeros(9, dtype=np.uint64)
        self.sample = None
        self.accumulation = None

    def render(self, inv_view_proj, eye, frame_count, albedo, normal, depth):
        h, w = depth.shape
        un = OracleDeferred((C.c_float * 16)(*np.asarray(inv_view_proj, dtype=np.float32).reshape(16)),
                            (C.c_float * 4)(float(eye[0]), float(eye[1]), float(eye[2]), 1.0), w, h, frame_count, (C.c_float * 40)(*self.sky40))
        a = np.ascontiguousarray(albedo, dtype=np.float32)
        n = np.ascontiguousarray(normal, dtype=np.float32)
        d = np.ascontiguousarray(depth, dtype=np.float32)
        self.sample = np.zeros((h, w, 3), dtype=np.float32)
        oracle().oracle_deferred_lighting(C.byref(self.scene), C.byref(un), _ptr(a), _ptr(n), _ptr(d), _ptr(self.sample), _ptr(self.counters), self.threads)
        if self.accumulation is None or self.accumulation.shape != self.sample.shape:
            self.accumulation = np.zeros_like(self.sample)
        oracle().oracle_deferred_resolve(_ptr(self.sample), _ptr(self.accumulation), w * h, frame_count)

    def stats(self) -> dict:
        return {k: int(v) for k, v in zip(COUNTER_NAMES, self.counters)}

    def display(self, exposure: float) -> np.ndarray:
        h, w, _ = self.accumulation.shape
        rgba = np.zeros((h, w, 4), dtype=np.float32)
        rgba[..., :3] = self.accumulation
        out = np.zeros((h, w), dtype=np.uint32)
        oracle().oracle_display(_ptr(rgba), w * h, 1.0, exposure, _ptr(out))
        return out


def rmse(a: np.ndarray, b: np.ndarray) -> float:
    """sqrt(mean over pixels x 3 channels of (delta)^2) on linear HDR; NaN == NaN counts as equal."""
    a3, b3 = a[..., :3].astype(np.float64), b[..., :3].astype(np.float64)
    d = a3 - b3
    both_nan = np.isnan(a3) & np.isnan(b3)
    d[both_nan] = 0.0
    same_inf = np.isinf(a3) & (a3 == b3)
    d[same_inf] = 0.0
    return float(np.sqrt(np.mean(d * d)))

"""The deferred renderer's lighting + resolve passes (csrc/deferred.cuh) against their restatement in
oracle/oracle.cpp (pt/deferred_renderer_lighting_pass.wgsl:96-186, pt/deferred_renderer_resolve_pass.wgsl:34-53).

The G-buffer is an input of the pass.  The tests build one the way the reference's rasteriser would fill it for the same
view — reverse-Z depth of the first surface along each pixel's view ray, a unit normal, an albedo — plus sky texels
(depth 0).  Control flow uses the same strict-fp32 operations on both sides, so the ray counters are equal and the
sample buffer agrees to rounding of the terminal sky/pow terms (RMSE tolerance 1e-4, as for the path tracer)."""
import numpy as np
import pytest

import _oracle as O
import rayfinder_b200 as rf

pytestmark = pytest.mark.gpu
RMSE_TOLERANCE = 1e-4


def look_at(eye, centre, up):
    f = centre - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m


def reverse_z_perspective(vfov, aspect, near, far):
    t = 1.0 / np.tan(vfov / 2.0)
    m = np.zeros((4, 4))
    m[0, 0], m[1, 1] = t / aspect, t
    m[2, 2], m[2, 3] = near / (far - near), far * near / (far - near)  # depth 1 at near, 0 at far
    m[3, 2] = -1.0
    return m


def make_gbuffer(pt, w, h, eye, centre, seed):
    """(inverse view-projection with m[c] = column c, albedo, encoded normal, reverse-Z depth) for a pinhole view."""
    rng = np.random.default_rng(seed)
    vp = reverse_z_perspective(np.radians(70.0), w / h, 0.05, 500.0) @ look_at(eye, centre, np.array([0.0, 1.0, 0.0]))
    inv = np.linalg.inv(vp)
    px, py = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5)
    ndc = np.stack([2 * px / w - 1, 2 * (1 - py / h) - 1, np.full_like(px, 0.5), np.ones_like(px)], axis=-1)
    world = ndc @ inv.T
    world = world[..., :3] / world[..., 3:]
    d = world - eye
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.concatenate([np.broadcast_to(eye, d.shape), d], axis=-1).reshape(-1, 6).astype(np.float32)
    hit, p_t, _ = O.oracle_intersect(pt.bvh_nodes, O.triangles9(pt), rays, 1000.0)
    hit = hit.reshape(h, w)
    t = p_t[:, 3].reshape(h, w).astype(np.float64)
    surface = eye + d * t[..., None]
    clip = np.concatenate([surface, np.ones((h, w, 1))], axis=-1) @ vp.T
    with np.errstate(divide="ignore", invalid="ignore"):
        depth = np.where(hit, clip[..., 2] / clip[..., 3], 0.0).astype(np.float32)
    assert (depth[hit] > 0).all()
    normal = np.zeros((h, w, 4), dtype=np.float32)
    n = -d + rng.normal(scale=0.3, size=d.shape)  # roughly facing the viewer
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    normal[..., :3] = (0.5 * n + 0.5).astype(np.float32)
    albedo = np.zeros((h, w, 4), dtype=np.float32)
    albedo[..., :3] = rng.uniform(0.1, 0.9, size=(h, w, 3))
    albedo[..., 3] = 1.0
    return inv.T.astype(np.float32).copy(), albedo, normal, depth  # inv.T rows = columns of the matrix


@pytest.mark.parametrize("scene,w,h", [("duck", 160, 120), ("sponza", 320, 180)])
def test_deferred_lighting_matches_oracle(duck_pt, sponza_pt, scene, w, h):
    pt = duck_pt if scene == "duck" else sponza_pt
    if scene == "duck":
        lo, hi = pt.bvh_nodes["aabb_min"][0].astype(np.float64), pt.bvh_nodes["aabb_max"][0].astype(np.float64)
        centre = 0.5 * (lo + hi)
        eye = centre + np.array([0.9, 0.5, 1.1]) * (hi - lo).max()
    else:
        eye, centre = np.array([1.22, 1.25, -1.25]), np.array([-5.0, 0.5, 6.0])
    sky = rf.Sky(turbidity=2.0, sun_zenith_degrees=35.0, sun_azimuth_degrees=20.0)
    cam = rf.fly_camera(w, h)
    ren = rf.ReferencePathTracer(rf.RenderParameters((w, h), cam, rf.SamplingParams(1, 2)), (w, h), rf.SceneArrays.from_pt(pt))
    orc = O.OracleDeferredLighting(pt, rf.sky_state(sky))
    inv, albedo, normal, depth = make_gbuffer(pt, w, h, eye, centre, seed=7)
    assert 0.05 < (depth == 0).mean() < 0.95 or scene == "sponza"
    for frame in (0, 1, 2, 1048577):  # 0 restarts the moving average; 2^20 + 1 wraps the blue-noise cycle
        ren.reset_stats()
        orc.counters[:] = 0
        ren.render_deferred_lighting(inv, eye, frame, albedo, normal, depth, sky=sky, exposure=0.5)
        orc.render(inv, eye, frame, albedo, normal, depth)
        sample, accumulation, display = ren.read_deferred()
        stats, expected = ren.stats(), orc.stats()
        for key in O.COUNTER_NAMES:
            assert stats[key] == expected[key], (frame, key)
        assert stats["paths"] == int((depth != 0).sum()) and stats["shadow_rays"] >= stats["paths"]
        assert np.isfinite(sample).all()
        assert O.rmse(sample, orc.sample) < RMSE_TOLERANCE, (frame, O.rmse(sample, orc.sample))
        assert O.rmse(accumulation, orc.accumulation) < RMSE_TOLERANCE
        d8 = np.abs(display.view(np.uint8).astype(np.int16) - orc.display(0.5).view(np.uint8).astype(np.int16))
        assert d8.max() <= 1
    # the sky texels carry the solar disk term where the view ray points at the sun
    assert sample[depth == 0].size == 0 or sample[depth == 0].min() > 0.0
    ren.close()


def test_deferred_argument_validation(duck_pt):
    w, h = 32, 16
    ren = rf.ReferencePathTracer(rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, 2)), (w, h), rf.SceneArrays.from_pt(duck_pt))
    big = np.zeros((h + 1, w), dtype=np.float32)
    with pytest.raises(rf.RayfinderError):
        ren.render_deferred_lighting(np.eye(4), (0, 0, 0), 0, np.zeros((h + 1, w, 4)), np.zeros((h + 1, w, 4)), big)
    with pytest.raises(rf.RayfinderError):
        ren._deferred_size = (w, h)
        ren.read_deferred()  # nothing rendered yet
    ren.close()

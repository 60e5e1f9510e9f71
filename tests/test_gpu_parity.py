"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle on the same inputs.

Bars (BASELINE.json north_star): BVH node-visit counts bit-exact vs the bvh-visualizer traversal; pixel
radiance RMSE < 1e-4 on linear HDR vs the renderer restatement.  Everything except the terminal sky term
(libm vs libdevice transcendentals) is in fact bit-exact, which the tests also assert through the
per-frame work counters and — on sun-only frames — through exact image equality.
"""
import numpy as np
import pytest

import _oracle as O
import rayfinder_b200 as rf
from rayfinder_b200 import capi

pytestmark = pytest.mark.gpu

RMSE_TOLERANCE = 1e-4  # fp32 tolerance stated by BASELINE.json:north_star


def make_renderer(pt, w, h, cam, spp, bounces, sky=None, max_size=None, exposure=1.0):
    params = rf.RenderParameters((w, h), cam, rf.SamplingParams(spp, bounces), sky or rf.Sky(), exposure)
    return rf.ReferencePathTracer(params, max_size or (w, h), rf.SceneArrays.from_pt(pt)), params


def assert_counters_equal(gpu_stats, oracle_stats):
    for key in O.COUNTER_NAMES:
        assert gpu_stats[key] == oracle_stats[key], key


# ---- config 1: node-count image ------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", [1, 2])
def test_duck_node_counts_bit_exact(duck_pt, golden, kernel):
    g = golden["ref_duck_node_counts"]
    scene = rf.TraversalScene(duck_pt.bvh_nodes, O.triangles9(duck_pt))
    scene.set_kernel(kernel)
    for (w, h) in ((512, 512), (1280, 720)):
        cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, w, h)
        assert np.array_equal(rf.camera_to_array(cam).view(np.uint32), g[f"camera_{w}x{h}"].view(np.uint32))
        counts, ms = scene.bvh_visualizer_node_counts(cam, w, h)
        assert np.array_equal(counts, g[f"counts_{w}x{h}"].astype(np.uint32))
        assert ms > 0
    # ragged sizes (not multiples of the 8x4 warp blocks) against the oracle
    for (w, h) in ((1, 1), (7, 3), (33, 17), (130, 67)):
        cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, w, h)
        counts, _ = scene.bvh_visualizer_node_counts(cam, w, h)
        expected, _ = O.oracle_node_counts(duck_pt.bvh_nodes, O.triangles9(duck_pt), rf.camera_to_array(cam), w, h, rf.FLT_MAX)
        assert np.array_equal(counts, expected)


def test_sponza_node_counts_bit_exact(sponza_pt):
    """Interior benchmark camera, full 1080p: every per-pixel count equals the oracle's, with either traversal kernel."""
    w, h = 1920, 1080
    tris = O.triangles9(sponza_pt)
    scene = rf.TraversalScene(sponza_pt.bvh_nodes, tris)
    cam = rf.fly_camera(w, h)
    counts, _ = scene.bvh_visualizer_node_counts(cam, w, h)
    expected, _ = O.oracle_node_counts(sponza_pt.bvh_nodes, tris, rf.camera_to_array(cam), w, h, rf.FLT_MAX)
    assert np.array_equal(counts, expected)
    scene.set_kernel(2)
    counts2, _ = scene.bvh_visualizer_node_counts(cam, w, h)
    assert np.array_equal(counts2, expected)
    assert 85 < counts.mean() < 95 and counts.max() < 1000  # SURVEY.md §6 probe: mean 89.7, max 369


@pytest.mark.parametrize("kernel", [1, 2])
def test_ray_intersect_bvh_batch_bit_exact(duck_pt, golden, kernel):
    """The reference's tests/bvh.cpp grid + random/axis-parallel rays: hit, p, t, nodesVisited bit-exact."""
    g = golden["ref_duck_bvh_test"]
    tris = O.triangles9(duck_pt)
    scene = rf.TraversalScene(duck_pt.bvh_nodes, tris)
    scene.set_kernel(kernel)
    hit, p_t, visited = scene.ray_intersect_bvh(g["rays"].reshape(-1, 6), 1000.0)
    assert np.array_equal(hit, g["hit"])
    assert np.array_equal(p_t.view(np.uint32), g["p_t"].view(np.uint32))
    assert np.array_equal(visited, g["nodes_visited"])

    rng = np.random.default_rng(5)
    lo, hi = duck_pt.bvh_nodes["aabb_min"][0], duck_pt.bvh_nodes["aabb_max"][0]
    n = 100_003
    origin = rng.uniform(lo - 1.0, hi + 1.0, (n, 3))
    d = rng.uniform(lo, hi, (n, 3)) - origin
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:600] = np.eye(3)[rng.integers(0, 3, 600)] * rng.choice([-1.0, 1.0], (600, 1))  # invDir = +-inf, 0*inf = NaN
    d[600:700] *= rng.uniform(0.1, 10.0, (100, 1))                                    # unnormalised directions
    rays = np.concatenate([origin, d], axis=1).astype(np.float32)
    for t_max in (rf.FLT_MAX, 1.25, 1e-3):
        a = scene.ray_intersect_bvh(rays, t_max)
        b = O.oracle_intersect(duck_pt.bvh_nodes, tris, rays, t_max)
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert np.array_equal(a[2], b[2])
    # NaN hazards: axis-parallel rays whose origin lies exactly on a slab plane of some BVH node (0 * inf = NaN in
    # the reference's slab products; std::max/std::min operand order decides what survives)
    nodes = duck_pt.bvh_nodes
    pick = rng.integers(0, nodes.size, 4000)
    origin = rng.uniform(lo - 0.5, hi + 0.5, (4000, 3)).astype(np.float32)
    direction = rng.normal(size=(4000, 3)).astype(np.float32)
    for k in range(4000):
        axis = k % 3
        origin[k, axis] = nodes["aabb_min" if (k // 3) % 2 else "aabb_max"][pick[k]][axis]
        direction[k, axis] = 0.0 if (k // 6) % 2 else -0.0
        if k % 5 == 0:  # two zero components, second origin coordinate on a plane of the root box
            other = (axis + 1) % 3
            direction[k, other] = 0.0
            origin[k, other] = nodes["aabb_min"][0][other]
    hazard = np.concatenate([origin, direction], axis=1)
    with np.errstate(all="ignore"):
        a = scene.ray_intersect_bvh(hazard, rf.FLT_MAX)
        b = O.oracle_intersect(nodes, tris, hazard, rf.FLT_MAX)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    # rays the fast path must not take at all: NaN / inf direction components, non-finite origins
    weird = rays[:64].copy()
    weird[0:16, 3] = np.inf
    weird[16:32, 4] = np.nan
    weird[32:48, 0] = np.inf
    weird[48:64, 2] = np.nan
    with np.errstate(all="ignore"):
        a = scene.ray_intersect_bvh(weird, 100.0)
        b = O.oracle_intersect(nodes, tris, weird, 100.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    # empty batch
    hit, p_t, visited = scene.ray_intersect_bvh(np.zeros((0, 6), np.float32), 1.0)
    assert hit.size == 0 and visited.size == 0


def test_aabb_and_triangle_known_answers_on_device():
    """reference tests/aabb.cpp:61-132 and tests/intersection.cpp:9-28 through the traversal kernel: a
    one-leaf BVH whose box is the test box."""
    cases = [((-2, 0, 0), (1, 0, 0), (-1, -1, -1), (1, 1, 1), True), ((0, -1, 0), (0, 1, 0), (-1, 0, -1), (1, 1, 1), True),
             ((0, 0, -1), (0, 0, 1), (-1, -1, 0), (1, 1, 1), True), ((-1, -1, -1), (1, 1, 1), (-1, -1, -1), (1, 1, 1), True),
             ((-2, 0, -1), (0, 1, 0), (-1, -1, -1), (1, 1, 1), False)]
    for origin, direction, lo, hi, expected in cases:
        node = np.zeros(1, dtype=rf.BVH_NODE_DTYPE)
        node["aabb_min"], node["aabb_max"] = lo, hi
        node["triangle_count"], node["split_axis"] = 1, 0xFFFFFFFF
        # a big triangle through the box centre facing the ray, so "box hit" <=> "triangle tested and hit"
        c = (np.array(lo, np.float32) + np.array(hi, np.float32)) / 2
        dvec = np.array(direction, np.float32)
        a = np.cross(dvec, [0.3, 0.5, 0.7]); a /= np.linalg.norm(a)
        b = np.cross(dvec, a)
        tri = np.concatenate([c - 50 * a - 50 * b, c + 50 * a - 50 * b, c + 50 * b]).astype(np.float32)[None]
        scene = rf.TraversalScene(node, tri)
        ray = np.array([[*origin, *direction]], dtype=np.float32)
        hit, _, visited = scene.ray_intersect_bvh(ray, 100.0)
        oh, _, ov = O.oracle_intersect(node, tri, ray, 100.0)
        assert bool(hit[0]) == bool(oh[0]) == expected and visited[0] == ov[0] == 1
    node = np.zeros(1, dtype=rf.BVH_NODE_DTYPE)
    node["aabb_min"], node["aabb_max"], node["triangle_count"], node["split_axis"] = (0, 0, 1), (1, 1, 1), 1, 0xFFFFFFFF
    tri = np.array([[0, 0, 1, 1, 0, 1, 0, 1, 1]], dtype=np.float32)
    hit, p_t, _ = rf.TraversalScene(node, tri).ray_intersect_bvh(np.array([[0, 0, 0, 0, 0, 1]], np.float32), 1000.0)
    assert hit[0] and abs(p_t[0, 0]) < 1e-3 and abs(p_t[0, 1]) < 1e-3 and p_t[0, 2] == pytest.approx(1.0, rel=1e-3)


# ---- configs 2/3: radiance ---------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,spp,bounces", [(96, 64, 2, 4), (200, 120, 1, 8), (65, 33, 3, 1)])
def test_duck_radiance_matches_oracle(duck_pt, golden, w, h, spp, bounces):
    cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, w, h)
    ren, _ = make_renderer(duck_pt, w, h, cam, spp, bounces)
    orc = O.OracleRenderer(duck_pt, w, h, rf.camera_to_array(cam), rf.sky_state(rf.Sky()), spp, bounces)
    for _ in range(spp):
        ren.render()
        orc.render()
    img, acc = ren.read_hdr()
    assert acc == spp == orc.accumulated
    assert_counters_equal(ren.stats(), orc.stats())
    assert O.rmse(img, orc.image) < RMSE_TOLERANCE
    assert np.all(img[..., 3] == 0)
    if (w, h, spp, bounces) == (96, 64, 2, 4):  # committed fixture
        g = golden["oracle_duck_hdr"]
        assert O.rmse(img, g["image"]) < RMSE_TOLERANCE
        assert ren.stats()["closest_rays"] == int(g["counters"][1])


def test_thin_lens_and_sky_parameters(duck_pt):
    """Aperture > 0 (pointInUnitDisk path), non-default sky, turbidity interpolation."""
    w, h = 128, 96
    lo, hi = duck_pt.bvh_nodes["aabb_min"][0], duck_pt.bvh_nodes["aabb_max"][0]
    centre = (lo + hi) / 2
    cam = rf.create_camera(centre + np.array([1.5, 0.8, 2.0], np.float32), centre, 0.08, 2.4, rf.degrees_to_radians(55.0), w / h)
    sky = rf.Sky(turbidity=3.5, albedo=(0.3, 0.6, 0.1), sun_zenith_degrees=62.0, sun_azimuth_degrees=140.0)
    ren, _ = make_renderer(duck_pt, w, h, cam, 4, 3, sky=sky)
    orc = O.OracleRenderer(duck_pt, w, h, rf.camera_to_array(cam), rf.sky_state(sky), 4, 3)
    for _ in range(4):
        ren.render()
        orc.render()
    img, _ = ren.read_hdr()
    assert_counters_equal(ren.stats(), orc.stats())
    assert O.rmse(img, orc.image) < RMSE_TOLERANCE


def test_sponza_radiance_matches_oracle(sponza_pt):
    """BASELINE.json configs[1] at a resolution the oracle finishes in seconds (same camera, 8 bounces)."""
    w, h, bounces = 480, 270, 8
    cam = rf.fly_camera(w, h)
    ren, _ = make_renderer(sponza_pt, w, h, cam, 1, bounces)
    orc = O.OracleRenderer(sponza_pt, w, h, rf.camera_to_array(cam), rf.sky_state(rf.Sky()), 1, bounces)
    ren.render()
    orc.render()
    img, _ = ren.read_hdr()
    assert_counters_equal(ren.stats(), orc.stats())
    err = O.rmse(img, orc.image)
    assert err < RMSE_TOLERANCE, err
    assert img[..., :3].max() > 1.0 and np.isfinite(img).all()


def test_sponza_accumulation_and_frame_counter(sponza_pt):
    """configs[2] semantics at reduced size: N frames accumulate into the sum buffer, frameCount keeps
    running after convergence, further render() calls change nothing (fsMain:51), a parameter change resets."""
    w, h, spp, bounces = 160, 90, 6, 4
    cam = rf.fly_camera(w, h)
    ren, params = make_renderer(sponza_pt, w, h, cam, spp, bounces)
    orc = O.OracleRenderer(sponza_pt, w, h, rf.camera_to_array(cam), rf.sky_state(rf.Sky()), spp, bounces)
    ren.set_frame_count(3)  # the sequence index is frameCount % spp (wgsl:607)
    orc.frame_count = 3
    for k in range(spp):
        ren.render()
        orc.render()
        assert ren.render_progress_percentage() == pytest.approx(100.0 * (k + 1) / spp)
    img, acc = ren.read_hdr()
    assert acc == spp and ren.frame_count == 3 + spp
    assert_counters_equal(ren.stats(), orc.stats())
    assert O.rmse(img, orc.image) < RMSE_TOLERANCE
    ren.render()  # converged: nothing traced, frame counter still advances
    img2, acc2 = ren.read_hdr()
    assert acc2 == spp and ren.frame_count == 4 + spp and np.array_equal(img.view(np.uint32), img2.view(np.uint32))
    assert ren.stats()["paths"] == spp * w * h
    # convergence: mean of 6 samples is closer to the mean of 24 than a single sample is
    params.exposure = 2.0  # any change resets the accumulation (reference_path_tracer.cpp:556-563)
    ren.set_render_parameters(params)
    assert ren.accumulated_sample_count == 0 and ren.render_progress_percentage() == 0.0
    ren.render()
    img3, acc3 = ren.read_hdr()
    assert acc3 == 1
    # display transform (fsMain:59-63) within one 8-bit step of the oracle's
    orc2 = O.OracleRenderer(sponza_pt, w, h, rf.camera_to_array(cam), rf.sky_state(rf.Sky()), spp, bounces)
    orc2.frame_count = ren.frame_count - 1
    orc2.render()
    assert O.rmse(img3, orc2.image) < RMSE_TOLERANCE
    disp = ren.read_display()
    expected = orc2.display(2.0)
    diff = np.abs((disp[..., None] >> np.array([0, 8, 16, 24]) & 0xFF).astype(int) - (expected[..., None] >> np.array([0, 8, 16, 24]) & 0xFF).astype(int))
    assert diff.max() <= 1 and np.all(disp >> 24 == 255)


def test_sponza_full_size_properties(sponza_pt):
    """configs[1] at BASELINE's full size, through size-independent properties: determinism (two renderers,
    bit-identical images), tile-partition linearity (sum over 8 disjoint rank images == single image, bit for
    bit — configs[3]'s reduce), bounded path lengths, and per-frame counter identities."""
    w, h, bounces = 1920, 1080, 8
    cam = rf.fly_camera(w, h)
    ren, _ = make_renderer(sponza_pt, w, h, cam, 1, bounces)
    ren.render()
    full, _ = ren.read_hdr()
    s = ren.stats()
    assert s["paths"] == w * h and s["closest_rays"] >= s["paths"]
    assert s["shadow_rays"] <= s["closest_rays"] <= bounces * s["paths"]
    assert s["closest_nodes_visited"] >= s["closest_rays"] and s["shadow_nodes_visited"] >= s["shadow_rays"]
    assert np.isfinite(full).all()

    ren.set_tile_partition(0, 1)  # same renderer again: deterministic
    ren.render()
    again, _ = ren.read_hdr()
    assert np.array_equal(full.view(np.uint32), again.view(np.uint32))

    from rayfinder_b200 import distributed as rfd
    world = 8
    owner = rfd.tile_owner(w, h, world)
    total = np.zeros_like(full)
    rays = 0
    for rank in range(world):
        ren.reset_stats()
        ren.set_tile_partition(rank, world)
        ren.render()
        part, _ = ren.read_hdr()
        assert np.all(part[owner != rank] == 0.0)
        assert ren.stats()["paths"] == rfd.owned_pixel_count(w, h, rank, world)
        rays += ren.stats()["closest_rays"]
        total += part
    assert np.array_equal(total.view(np.uint32), full.view(np.uint32))
    assert rays == s["closest_rays"]


def test_framebuffer_resize_within_max(duck_pt):
    """maxFramebufferSize allocation (reference_path_tracer.cpp:186-190): smaller framebuffers reuse it."""
    cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, 64, 48)
    ren, params = make_renderer(duck_pt, 64, 48, cam, 1, 2, max_size=(256, 256))
    ren.render()
    small, _ = ren.read_hdr()
    params2 = rf.RenderParameters((250, 130), rf.bvh_visualizer_camera(duck_pt.bvh_nodes, 250, 130), rf.SamplingParams(1, 2))
    ren.set_render_parameters(params2)
    ren.render()
    big, _ = ren.read_hdr()
    orc = O.OracleRenderer(duck_pt, 250, 130, rf.camera_to_array(params2.camera), rf.sky_state(rf.Sky()), 1, 2)
    orc.frame_count = 1
    orc.render()
    assert big.shape == (130, 250, 4) and O.rmse(big, orc.image) < RMSE_TOLERANCE
    with pytest.raises(rf.RayfinderError):
        ren.set_render_parameters(rf.RenderParameters((300, 100), cam, rf.SamplingParams(1, 2)))
    assert ren.average_renderpass_duration_ms() >= 0.0


def test_texture_limit_error_message(duck_pt):
    """reference_path_tracer.cpp:256-263: texture data above the 1 GiB binding limit is rejected."""
    scene = rf.SceneArrays.from_pt(duck_pt)
    huge = np.zeros((16385, 16384), dtype=np.uint32)  # > 1 GiB of BGRA8
    scene = rf.SceneArrays(scene.bvh_nodes, scene.position_attributes, scene.vertex_attributes, [huge])
    cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, 8, 8)
    with pytest.raises(rf.RayfinderError) as e:
        rf.ReferencePathTracer(rf.RenderParameters((8, 8), cam, rf.SamplingParams(1, 1)), (8, 8), scene)
    assert str(e.value) == f"Texture buffer size ({huge.size * 4}) exceeds maxStorageBufferBindingSize ({1 << 30})."


@pytest.mark.parametrize("kernel,sub_frames,persistent,variant,block,tri_min,refill_min,evict_max", [
    (1, 1, 0, 3, 256, 4, 4, 0), (1, 2, 0, 3, 256, 4, 4, -1), (1, 4, 0, 2, 64, 1, 1, 8), (1, 3, 0, 1, 128, 32, 32, 32), (1, 1, 0, 11, 256, 8, 16, 1),
    (1, 1, 0, 3, 256, 4, 4, 32), (1, 2, 0, 3, 256, 4, 4, 0), (1, 1, 1, 3, 256, 4, 4, -1), (1, 2, 1, 3, 128, 2, 8, -1),
    (2, 1, 0, 3, 256, 4, 4, 0), (2, 2, 0, 7, 256, 4, 4, 0), (2, 3, 0, 1, 256, 1, 1, 0), (2, 1, 0, 5, 256, 32, 32, 0), (2, 4, 0, 7, 256, 8, 2, 0)])
def test_results_do_not_depend_on_scheduling(duck_pt, kernel, sub_frames, persistent, variant, block, tri_min, refill_min, evict_max):
    """The traversal kernel (child-pair records or one node per visit), sub-frame pipelining, the experimental persistent
    kernel, the compile-time scheduling variants, block sizes, the run-time knobs and the straggler hand-over (rays moved to
    another lane in mid-traversal) change how warps are kept busy — never a counter or a pixel."""
    w, h, spp, bounces = 150, 70, 2, 5
    cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, w, h)
    ren, _ = make_renderer(duck_pt, w, h, cam, spp, bounces)
    ren.set_option("trace_kernel", 1)
    ren.set_pipeline(1, 0, 3, 256)
    ren.set_tail_policy(0)
    ren.render(), ren.render()
    ref_img, _ = ren.read_hdr()
    ref_stats = ren.stats()
    assert ref_stats["trace_kernel"] == 1 and ref_stats["node_records_loaded"] == ref_stats["closest_nodes_visited"] + ref_stats["shadow_nodes_visited"]
    ren2, _ = make_renderer(duck_pt, w, h, cam, spp, bounces)
    ren2.set_option("trace_kernel", kernel)
    ren2.set_tuning(tri_min, refill_min, 4)
    if kernel == 2:
        ren2.set_option("pair_variant", variant)
        ren2.set_pipeline(sub_frames, persistent, -1, 0)
    else:
        ren2.set_pipeline(sub_frames, persistent, variant, block)
    ren2.set_tail_policy(evict_max)
    ren2.render(), ren2.render()
    img, _ = ren2.read_hdr()
    stats = ren2.stats()
    assert stats["trace_kernel"] == (1 if persistent else kernel)
    if kernel == 2:  # every record decides two visits; a ray that ends early leaves pushed entries unvisited
        visits = stats["closest_nodes_visited"] + stats["shadow_nodes_visited"]
        rays = stats["closest_rays"] + stats["shadow_rays"]
        assert (stats["closest_nodes_visited"] - stats["closest_rays"]) // 2 <= stats["node_records_loaded"] <= visits - rays
    for key in O.COUNTER_NAMES:
        assert stats[key] == ref_stats[key], key
    assert np.array_equal(img.view(np.uint32), ref_img.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("sub_frames,evict_max", [(1, 0), (2, 0), (2, 8)])
def test_sponza_walk_in_place_is_exact(sponza_pt, sub_frames, evict_max):
    """Option walk_in_place: once a staged traversal launch's queue is dry, a warp left with one ray walks it with all 32 lanes in
    place (the ray moves from its lane to the warp in mid-traversal, through registers and the warp's own stack memory)."""
    w, h, bounces = 480, 270, 8
    cam = rf.fly_camera(w, h)
    images, stats = [], []
    for walk in (0, 1):
        ren, _ = make_renderer(sponza_pt, w, h, cam, 2, bounces)
        ren.set_option("trace_kernel", 1)
        ren.set_pipeline(sub_frames, 0, 3, 256)
        ren.set_tail_policy(evict_max)
        ren.set_option("walk_in_place", walk)
        ren.render(), ren.render()
        images.append(ren.read_hdr()[0])
        stats.append(ren.stats())
        ren.close()
    for key in O.COUNTER_NAMES:
        assert stats[0][key] == stats[1][key], key
    assert np.array_equal(images[0].view(np.uint32), images[1].view(np.uint32))


@pytest.mark.parametrize("evict_max,sub_frames,window_mode", [(1, 1, 0), (8, 1, 0), (32, 2, 0), (-1, -1, 0), (8, 1, 1), (32, 2, 1), (-1, -1, 1)])
def test_sponza_tail_hand_over_is_exact(sponza_pt, evict_max, sub_frames, window_mode):
    """The warp-per-ray tail kernel (csrc/straggler.cuh: 32-node windows, lane-parallel slab and first-triangle tests,
    one triangle per lane in multi-triangle leaves) resumes rays in mid-traversal.  On Sponza's long grazing rays —
    including the 1/256 of the shadow rays that are axis-parallel and take the NaN re-test — every counter and
    every pixel equals the run in which each ray ends on the lane it started on."""
    w, h, bounces = 480, 270, 8
    cam = rf.fly_camera(w, h)
    ren, _ = make_renderer(sponza_pt, w, h, cam, 2, bounces)
    ren.set_option("trace_kernel", 1)  # the tail hand-over belongs to the one-node-per-visit kernel
    ren.set_pipeline(1, 0, 3, 256)
    ren.set_tail_policy(0)
    ren.render(), ren.render()
    ref_img, _ = ren.read_hdr()
    ref_stats = ren.stats()
    assert ref_stats["evict_max"] == 0 and ref_stats["sub_frames"] == 1
    ren2, _ = make_renderer(sponza_pt, w, h, cam, 2, bounces)
    ren2.set_option("trace_kernel", 1)
    ren2.set_option("tail_window_mode", window_mode)  # 1: windows staged by cp.async.bulk + mbarrier, next window copied ahead
    ren2.set_pipeline(sub_frames, 0, 3, 256)
    ren2.set_tail_policy(evict_max)
    ren2.render(), ren2.render()
    img, _ = ren2.read_hdr()
    stats = ren2.stats()
    assert stats["evict_max"] == (0 if evict_max < 0 else evict_max)  # automatic: off — warps walk their last ray in place (walk_in_place)
    if evict_max > 0:
        assert stats["kernel_launches"] > ref_stats["kernel_launches"]
    for key in O.COUNTER_NAMES:
        assert stats[key] == ref_stats[key], key
    assert np.array_equal(img.view(np.uint32), ref_img.view(np.uint32))


@pytest.mark.parametrize("block,slots,tail_paths,sub_frames,shade_wait", [
    (512, 0, 0, -1, 16), (256, 0, 0, 1, 16), (256, 64, 1, 1, 0), (512, 96, 401, 2, 4), (256, 512, 2, 3, 64), (512, 1024, 31, 1, 16)])
def test_sponza_persistent_kernel_is_exact(sponza_pt, block, slots, tail_paths, sub_frames, shade_wait):
    """The frame as one persistent launch (csrc/mega.cuh): block-local path records, shared-memory ready / hit rings, a
    path's shadow ray and next closest-hit ray on one lane with the shadow result folded in by the shading warp, lagging
    paths first, and whole-warp walks of the last rays (rays moved from a lane to the warp in mid-traversal).  Block size,
    path slots per block (few slots: pixels are taken as paths end; many: the whole share starts at once), the tail
    threshold (off, always, early) and the batching wait change the schedule, never a counter or a pixel of the staged
    pipeline's frame — on the scene whose grazing rays and axis-parallel shadow rays exercise the tail and the NaN re-test."""
    w, h, bounces = 480, 270, 8
    cam = rf.fly_camera(w, h)
    ren, _ = make_renderer(sponza_pt, w, h, cam, 2, bounces)
    ren.set_option("trace_kernel", 1)
    ren.set_pipeline(1, 0, 3, 256)
    ren.set_tail_policy(0)
    ren.render(), ren.render()
    ref_img, _ = ren.read_hdr()
    ref_stats = ren.stats()
    assert ref_stats["persistent_kernel"] == 0 and ref_stats["sub_frames"] == 1
    ren2, _ = make_renderer(sponza_pt, w, h, cam, 2, bounces)
    if sub_frames < 0:
        assert ren2.stats()["persistent_kernel"] == 1  # the automatic schedule of a frame this small
    ren2.set_pipeline(sub_frames, 2 if sub_frames < 0 else 1, -1, 0)
    ren2.set_option("mega_block", block)
    ren2.set_option("mega_slots", slots)
    ren2.set_option("tail_paths", tail_paths)
    ren2.set_option("shade_wait", shade_wait)
    ren2.render(), ren2.render()
    img, _ = ren2.read_hdr()
    stats = ren2.stats()
    assert stats["persistent_kernel"] == 1 and stats["sub_frames"] == (1 if sub_frames < 0 else sub_frames)
    assert stats["kernel_launches"] == 2 * 2 * stats["sub_frames"]  # per frame and tile set: the persistent launch + the accumulation
    for key in O.COUNTER_NAMES:
        assert stats[key] == ref_stats[key], key
    assert np.array_equal(img.view(np.uint32), ref_img.view(np.uint32))


@pytest.mark.parametrize("bounces,w,h,spp", [(1, 150, 70, 1), (2, 97, 61, 3), (16, 64, 48, 2), (8, 33, 5, 1)])
def test_persistent_kernel_bounce_counts_and_ragged_frames(duck_pt, bounces, w, h, spp):
    """One bounce (every path ends with its first shadow ray: the result-less hit-ring entry), two, sixteen (more levels than
    the priority tells apart is not reached, but deep chains are), frames smaller than one pixel group per block, accumulation
    over several samples: the persistent kernel equals the staged pipeline bit for bit."""
    cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, w, h)
    images, stats = [], []
    for persistent in (0, 1):
        ren, _ = make_renderer(duck_pt, w, h, cam, spp, bounces)
        ren.set_option("trace_kernel", 1)
        ren.set_pipeline(1, persistent, 3, 256)
        ren.set_tail_policy(0)
        for _ in range(spp + 1):  # one call more than samples: the converged frame traces nothing
            ren.render()
        img, acc = ren.read_hdr()
        assert acc == spp
        images.append(img)
        stats.append(ren.stats())
        ren.close()
    assert stats[0]["persistent_kernel"] == 0 and stats[1]["persistent_kernel"] == 1
    for key in O.COUNTER_NAMES:
        assert stats[0][key] == stats[1][key], key
    assert stats[1]["paths"] == spp * w * h
    assert np.array_equal(images[0].view(np.uint32), images[1].view(np.uint32))


def test_click_to_focus_matches_reference_formula(duck_pt):
    """pt/main.cpp:198-226: ray through the cursor, rayIntersectBvh(..., 1000.f, ...), dot(hit.p - position, forward)."""
    f32 = np.float32
    tris = O.triangles9(duck_pt)
    scene = rf.TraversalScene(duck_pt.bvh_nodes, tris)
    w, h = 640, 480
    cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, w, h)
    c = rf.camera_to_array(cam)
    origin, lower_left, horizontal, vertical = c[0:3], c[3:6], c[6:9], c[9:12]
    forward = np.cross(c[12:15], c[15:18]).astype(f32)  # up x right (camera.cpp: up = cross(right, forward))
    forward /= np.linalg.norm(forward)
    hits = 0
    for x, y in [(320.0, 240.0), (100.5, 50.25), (639.9, 479.9), (0.0, 0.0), (333.0, 300.0), (-1.0, 5.0), (640.0, 10.0), (300.25, 220.0),
                 (345.5, 255.0), (310.0, 262.75), (290.0, 240.0), (330.0, 215.5)]:
        got = scene.pick_focus_distance(cam, origin, forward, x, y, w, h)
        if not (0.0 <= x < w and 0.0 <= y < h):
            assert got is None
            continue
        u, v = f32(x) / f32(w), f32(1.0) - f32(y) / f32(h)
        d = ((lower_left + horizontal * u) + vertical * v) - origin
        d = d * (f32(1.0) / np.sqrt(f32(f32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])))
        hit, p_t, _ = O.oracle_intersect(duck_pt.bvh_nodes, tris, np.concatenate([origin, d])[None, :].astype(f32), 1000.0)
        if not hit[0]:
            assert got is None
            continue
        hits += 1
        rel = p_t[0, :3] - origin
        expected = f32(f32(rel[0] * forward[0] + rel[1] * forward[1]) + rel[2] * forward[2])
        assert f32(got) == expected
    assert hits >= 3

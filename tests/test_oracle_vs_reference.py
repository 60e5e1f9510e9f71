"""Pin the oracle (oracle/oracle.cpp) against the reference: committed fixtures generated from the
reference's own compiled CPU translation units (tests/golden/make_golden.py), the reference's
known-answer tests, and — when oracle/_ref is present — live comparisons on fresh random rays."""
import numpy as np
import pytest

import _oracle as O
import rayfinder_b200 as rf


def test_node_counts_match_reference_bvh_visualizer(duck_pt, golden):
    """BASELINE.json configs[0]: Duck 512x512 node-count image, bit-exact vs the reference traversal."""
    g = golden["ref_duck_node_counts"]
    tris = O.triangles9(duck_pt)
    for (w, h) in ((512, 512), (1280, 720)):
        counts, _ = O.oracle_node_counts(duck_pt.bvh_nodes, tris, g[f"camera_{w}x{h}"], w, h, rf.FLT_MAX)
        assert np.array_equal(counts, g[f"counts_{w}x{h}"].astype(np.uint32))
    assert int(g["counts_512x512"].astype(np.uint64).sum()) == 4839510  # SURVEY.md §6 probe
    assert int(g["counts_512x512"].max()) == 167


def test_bvh_grid_matches_reference_and_brute_force(duck_pt, golden):
    """reference tests/bvh.cpp:34-102: BVH result == brute force (hit/miss and t), plus bit-exact (hit, p, t,
    nodesVisited) vs the reference's rayIntersectBvh on the same 64x64 grid, rayTMax = 1000."""
    g = golden["ref_duck_bvh_test"]
    tris = O.triangles9(duck_pt)
    rays = g["rays"].reshape(-1, 6)
    hit, p_t, visited = O.oracle_intersect(duck_pt.bvh_nodes, tris, rays, 1000.0)
    assert np.array_equal(hit, g["hit"])
    assert np.array_equal(p_t.view(np.uint32), g["p_t"].view(np.uint32))
    assert np.array_equal(visited, g["nodes_visited"])
    # brute force over all triangles with the oracle's triangle test, sequentially shrinking tmax
    out4 = np.zeros(4, dtype=np.float32)
    step = 97  # a subset keeps the CPU suite fast; the full grid is covered by the bit-exact check above
    for k in range(0, rays.shape[0], step):
        ray = np.ascontiguousarray(rays[k])
        tmax, did = 1000.0, False
        for t in tris:
            if O.oracle().oracle_ray_intersect_triangle(O._ptr(ray), O._ptr(np.ascontiguousarray(t)), tmax, O._ptr(out4)):
                tmax, did = float(out4[3]), True
        assert did == bool(hit[k])
        if did:
            assert tmax == pytest.approx(float(p_t[k, 3]))


AABB_CASES = [  # reference tests/aabb.cpp:61-132
    ((-2, 0, 0), (1, 0, 0), (-1, -1, -1), (1, 1, 1), True),
    ((0, -1, 0), (0, 1, 0), (-1, 0, -1), (1, 1, 1), True),
    ((0, 0, -1), (0, 0, 1), (-1, -1, 0), (1, 1, 1), True),
    ((-1, -1, -1), (1, 1, 1), (-1, -1, -1), (1, 1, 1), True),
    ((-2, 0, -1), (0, 1, 0), (-1, -1, -1), (1, 1, 1), False),
]


@pytest.mark.parametrize("origin,direction,lo,hi,expected", AABB_CASES)
def test_ray_aabb_truth_table(origin, direction, lo, hi, expected):
    ray = np.array([*origin, *direction], dtype=np.float32)
    box = np.array([*lo, *hi], dtype=np.float32)
    with np.errstate(all="ignore"):
        assert bool(O.oracle().oracle_ray_intersect_aabb(O._ptr(ray), O._ptr(box), 100.0)) is expected
    if O.have_ref():
        assert bool(O.ref().ref_ray_intersect_aabb(O._ptr(ray), O._ptr(box), 100.0)) is expected


def test_ray_triangle_known_answer():
    """reference tests/intersection.cpp:9-28."""
    ray = np.array([0, 0, 0, 0, 0, 1], dtype=np.float32)
    tri = np.array([0, 0, 1, 1, 0, 1, 0, 1, 1], dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    assert O.oracle().oracle_ray_intersect_triangle(O._ptr(ray), O._ptr(tri), 1000.0, O._ptr(out))
    assert abs(out[0]) < 1e-3 and abs(out[1]) < 1e-3 and out[2] == pytest.approx(1.0, rel=1e-3) and out[3] == pytest.approx(1.0)
    if O.have_ref():
        ref_out = np.zeros(4, dtype=np.float32)
        assert O.ref().ref_ray_intersect_triangle(O._ptr(ray), O._ptr(tri), 1000.0, O._ptr(ref_out))
        assert np.array_equal(out.view(np.uint32), ref_out.view(np.uint32))


def test_create_camera_matches_reference(golden):
    g = golden["ref_cameras"]
    for inp, expected in zip(g["inputs"], g["cameras"]):
        out = np.zeros(19, dtype=np.float32)
        vfov = rf.degrees_to_radians(float(inp[8]))
        O.oracle().oracle_create_camera(O._ptr(np.ascontiguousarray(inp[0:3])), O._ptr(np.ascontiguousarray(inp[3:6])),
                                        float(inp[6]), float(inp[7]), vfov, float(inp[9]), O._ptr(out))
        assert np.array_equal(out.view(np.uint32), expected.view(np.uint32))


def test_sky_radiance_matches_reference_model(golden):
    """Device-side skyRadiance (wgsl:248-275) is the sky part of sky_state_radiance (hw_skymodel.c:182-223)
    when gamma is outside the solar disk."""
    g = golden["ref_sky_states"]
    for st, rad in zip(g["states"], g["radiance"]):
        sky40 = np.zeros(40, dtype=np.float32)
        sky40[:33] = st
        k = 0
        for th in (0.1, 0.7, 1.5):
            for ga in (0.2, 1.0, 2.5):
                for ch in range(3):
                    got = O.oracle().oracle_sky_radiance(O._ptr(sky40), th, ga, ch)
                    assert got == pytest.approx(float(rad[k]), rel=2e-6)
                    k += 1


def test_solar_constants_known_answer():
    """fp32 SOLAR_COS_THETA_MAX = 0x1.fffeb4p-1, SOLAR_INV_PDF = 6.216817e-05."""
    out = np.zeros(2, dtype=np.float32)
    O.oracle().oracle_solar_constants(O._ptr(out))
    assert float(out[0]) == float.fromhex("0x1.fffeb4p-1")
    assert float(out[1]) == pytest.approx(6.216817e-05, rel=1e-6)
    assert float(np.float32(1.0) - out[0]) == pytest.approx(9.894371e-06, rel=1e-6)


def test_oracle_b_regression_pin(duck_pt, golden):
    """Oracle B reproduces its committed output (restatement regression pin; not a reference output)."""
    g = golden["oracle_duck_hdr"]
    orc = O.OracleRenderer(duck_pt, 96, 64, g["camera"], g["sky"], int(g["spp"]), int(g["bounces"]))
    orc.render()
    orc.render()
    assert np.array_equal(orc.counters, g["counters"])
    assert np.array_equal(orc.path_lengths, g["path_lengths"])
    assert O.rmse(orc.image, g["image"]) < 1e-6  # libm may differ between boxes by an ulp in the sky term


def test_oracle_c_regression_pin(duck_pt, golden):
    """Oracle C (deferred lighting + resolve passes) reproduces its committed output on the committed G-buffer
    (restatement regression pin; not a reference output)."""
    g = golden["oracle_duck_deferred"]
    h, w = g["depth"].shape
    albedo = np.zeros((h, w, 4), dtype=np.float32)
    normal = np.zeros((h, w, 4), dtype=np.float32)
    albedo[..., :3], normal[..., :3] = g["albedo"], g["normal"]
    orc = O.OracleDeferredLighting(duck_pt, g["sky"])
    for frame in (0, 5):
        orc.render(g["inv"], g["eye"], frame, albedo, normal, g["depth"])
        assert O.rmse(orc.sample, g[f"sample_{frame}"]) < 1e-6  # libm may differ between boxes by an ulp in the sky term
        assert O.rmse(orc.accumulation, g[f"accumulation_{frame}"]) < 1e-6
    assert np.array_equal(orc.counters, g["counters"])
    # frame 5 is 0.1 * sample + 0.9 * the frame-0 accumulation (deferred_renderer_resolve_pass.wgsl:44-46)
    expected = np.float32(0.1) * g["sample_5"] + np.float32(0.9) * g["accumulation_0"]
    assert np.array_equal(expected, g["accumulation_5"])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_oracle_matches_reference_on_random_rays(duck_pt):
    rng = np.random.default_rng(11)
    lo, hi = duck_pt.bvh_nodes["aabb_min"][0], duck_pt.bvh_nodes["aabb_max"][0]
    n = 20000
    origin = rng.uniform(lo - 1.0, hi + 1.0, (n, 3))
    target = rng.uniform(lo, hi, (n, 3))
    d = target - origin
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:64] = np.eye(3)[rng.integers(0, 3, 64)] * rng.choice([-1.0, 1.0], (64, 1))  # axis-parallel: invDir = +-inf
    rays = np.concatenate([origin, d], axis=1).astype(np.float32)
    # NaN hazards: zero direction component and the origin exactly on a slab plane of some node (0 * inf)
    nodes = duck_pt.bvh_nodes
    pick = rng.integers(0, nodes.size, 3000)
    for k in range(3000):
        axis = k % 3
        rays[100 + k, axis] = nodes["aabb_min" if (k // 3) % 2 else "aabb_max"][pick[k]][axis]
        rays[100 + k, 3 + axis] = 0.0 if (k // 6) % 2 else -0.0
    for t_max in (rf.FLT_MAX, 1.5):
        a = O.oracle_intersect(duck_pt.bvh_nodes, O.triangles9(duck_pt), rays, t_max)
        b = O.ref_intersect(duck_pt.bvh_nodes, O.triangles9(duck_pt), rays, t_max)
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert np.array_equal(a[2], b[2])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_oracle_matches_reference_sponza_primary_rays(sponza_pt):
    w, h = 240, 135
    cam = rf.camera_to_array(rf.fly_camera(w, h))
    tris = O.triangles9(sponza_pt)
    a, _ = O.oracle_node_counts(sponza_pt.bvh_nodes, tris, cam, w, h, rf.FLT_MAX)
    b, _ = O.ref_node_counts(sponza_pt.bvh_nodes, tris, cam, w, h, rf.FLT_MAX)
    assert np.array_equal(a, b)
    assert 60 < a.mean() < 120  # SURVEY.md §6 probe: 89.7 nodes / primary ray at 1080p

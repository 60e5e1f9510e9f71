"""Host-side pieces of the product library (no GPU): createCamera, sky state, BVH builder, .pt codec —
checked against fixtures generated from the reference's own compiled code, and against oracle/_ref live
when it is present."""
import ctypes as C

import numpy as np
import pytest

import _oracle as O
import rayfinder_b200 as rf
from rayfinder_b200 import capi


def test_create_camera_bit_exact(golden):
    g = golden["ref_cameras"]
    for inp, expected in zip(g["inputs"], g["cameras"]):
        cam = rf.create_camera(inp[0:3], inp[3:6], float(inp[6]), float(inp[7]), rf.degrees_to_radians(float(inp[8])), float(inp[9]))
        assert np.array_equal(rf.camera_to_array(cam).view(np.uint32), expected.view(np.uint32))


def test_sky_state_bit_exact(golden):
    """rf_sky_state_new == the reference's sky_state_new (hw_skymodel.c:141-180) for a parameter grid."""
    g = golden["ref_sky_states"]
    for p, st in zip(g["params"], g["states"]):
        sky = rf.Sky(float(p[0]), (float(p[1]), float(p[2]), float(p[3])), float(p[4]), 0.0)
        got = rf.sky_state(sky)
        assert np.array_equal(got[:33].view(np.uint32), st.view(np.uint32)), p
        assert np.all(got[33:36] == 0) and got[39] == 0


def test_sky_state_defaults_known_answer():
    """SURVEY.md §8(c) anchors for Sky{} defaults."""
    s = rf.sky_state(rf.Sky())
    assert s[27:30] == pytest.approx([9.67027664, 16.4482307, 27.4176598], rel=1e-7)
    assert s[30:33] == pytest.approx([796325.938, 503392.188, 234451.922], rel=1e-7)
    assert s[0:3] == pytest.approx([-1.08006394, -0.166454494, 2.705446], rel=1e-6)
    assert s[36:39] == pytest.approx([0.5, 0.8660254, 0.0], abs=1e-6)


@pytest.mark.parametrize("sky,field", [(rf.Sky(turbidity=0.5), "turbidity"), (rf.Sky(turbidity=10.5), "turbidity"),
                                       (rf.Sky(albedo=(1.5, 0, 0)), "albedo"), (rf.Sky(sun_zenith_degrees=100.0), "elevation")])
def test_sky_state_out_of_range(sky, field):
    with pytest.raises(rf.RayfinderError) as e:
        rf.sky_state(sky)
    assert e.value.status == capi.RF_ERROR_OUT_OF_RANGE and field in str(e.value)


def test_blue_noise_table_matches_reference():
    bn = O.blue_noise_rg8()
    assert bn.size == 128 * 128 * 2
    assert (int(bn[0]), int(bn[1]), int(bn[-1])) == (139, 121, 2)  # blue_noise.c:4, :1728
    if O.have_ref():
        w, h = C.c_uint64(), C.c_uint64()
        p = O.ref().ref_blue_noise(C.byref(w), C.byref(h))
        assert (w.value, h.value) == (128, 128)
        assert np.array_equal(np.ctypeslib.as_array(p, shape=(32768,)), bn)


def test_build_bvh_reproduces_baked_tree(duck_pt):
    """Rebuilding over the baked (already leaf-ordered) triangles is a fixed point of the builder, and the
    tree has the shape the survey probe measured with the reference's bvh.cpp (8383 nodes / 4212 triangles)."""
    tris = O.triangles9(duck_pt).reshape(-1, 3, 3)
    nodes, idx = rf.build_bvh(tris)
    assert nodes.size == 8383 and tris.shape[0] == 4212
    leaves = nodes[nodes["triangle_count"] > 0]
    assert leaves.size == 4192 and leaves["triangle_count"].max() == 4
    assert np.all(leaves["split_axis"] == 0xFFFFFFFF) and np.all(leaves["second_child_offset"] == 0)
    assert sorted(idx.tolist()) == list(range(tris.shape[0]))
    # every triangle lies inside its leaf's box
    reordered = rf.reorder_attributes(tris, idx)
    for leaf in leaves[:: 97]:
        t = reordered[leaf["triangles_offset"]: leaf["triangles_offset"] + leaf["triangle_count"]]
        assert np.all(t >= leaf["aabb_min"]) and np.all(t <= leaf["aabb_max"])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_build_bvh_bit_identical_to_reference(duck_pt):
    rng = np.random.default_rng(3)
    tris = O.triangles9(duck_pt)
    shuffled = np.ascontiguousarray(tris[rng.permutation(tris.shape[0])])
    for t in (tris, shuffled, shuffled[:2], shuffled[:1], np.repeat(shuffled[:1], 5, axis=0)):
        t = np.ascontiguousarray(t)
        h = O.ref().ref_bvh_build(O._ptr(t), t.shape[0])
        n = O.ref().ref_bvh_num_nodes(h)
        ref_nodes = np.zeros(n, dtype=rf.BVH_NODE_DTYPE)
        ref_idx = np.zeros(t.shape[0], dtype=np.uint64)
        O.ref().ref_bvh_copy(h, O._ptr(ref_nodes), O._ptr(ref_idx))
        O.ref().ref_bvh_free(h)
        nodes, idx = rf.build_bvh(t.reshape(-1, 3, 3))
        assert nodes.tobytes() == ref_nodes.tobytes()
        assert np.array_equal(idx, ref_idx)


def _assert_host_bvh_equals_reference(t):
    t = np.ascontiguousarray(t, dtype=np.float32).reshape(-1, 9)
    h = O.ref().ref_bvh_build(O._ptr(t), t.shape[0])
    n = O.ref().ref_bvh_num_nodes(h)
    ref_nodes = np.zeros(n, dtype=rf.BVH_NODE_DTYPE)
    ref_idx = np.zeros(t.shape[0], dtype=np.uint64)
    O.ref().ref_bvh_copy(h, O._ptr(ref_nodes), O._ptr(ref_idx))
    O.ref().ref_bvh_free(h)
    nodes, idx = rf.build_bvh(t.reshape(-1, 3, 3))
    assert nodes.tobytes() == ref_nodes.tobytes()
    assert np.array_equal(idx, ref_idx)
    return nodes


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_build_bvh_bit_identical_to_reference_on_soups():
    """The soups tests/test_gpu_bvh_build.py feeds the GPU builder, here against the reference's own bvh.cpp: random
    clouds, coarse grids with -0.0f / +0.0f mixed in and boxes flat at zero (the sequential min/max folds decide which
    zero a box keeps), more than 255 identical centroids (one leaf), and SAH-says-leaf above 255 primitives (forced
    splits).  The host builder is the GPU builder's yardstick, so it is pinned on the same cases."""
    for n in (1, 2, 3, 4, 7, 33, 257, 1000, 20000):
        rng = np.random.default_rng(n)
        centres = rng.uniform(-10, 10, size=(n, 1, 3))
        _assert_host_bvh_equals_reference(centres + rng.normal(scale=0.3, size=(n, 3, 3)))
    rng = np.random.default_rng(11)
    for n, grid in ((300, 2), (5000, 3), (20000, 8)):
        tris = rng.integers(-grid, grid + 1, size=(n, 3, 3)).astype(np.float32)
        neg = rng.random(size=tris.shape) < 0.5
        tris = np.where((tris == 0) & neg, np.float32(-0.0), tris).astype(np.float32)
        flat = rng.random(n) < 0.3
        tris[flat, :, 1] = np.where(rng.random((flat.sum(), 3)) < 0.5, np.float32(-0.0), np.float32(0.0))
        _assert_host_bvh_equals_reference(tris)
    rng = np.random.default_rng(3)
    same = np.tile(rng.normal(size=(1, 3, 3)), (700, 1, 1)).astype(np.float32)
    nodes = _assert_host_bvh_equals_reference(same)
    assert len(nodes) == 1 and nodes["triangle_count"][0] == 700
    big = (rng.normal(scale=100.0, size=(3000, 3, 3)) + rng.normal(scale=0.01, size=(3000, 1, 3))).astype(np.float32)
    nodes = _assert_host_bvh_equals_reference(big)
    assert nodes["triangle_count"].max() <= 255
    _assert_host_bvh_equals_reference(np.concatenate([same, big]))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_build_bvh_bit_identical_to_reference_on_sponza(sponza_pt):
    tris = O.triangles9(sponza_pt)
    _assert_host_bvh_equals_reference(tris[np.random.default_rng(5).permutation(len(tris))])


def test_pt_round_trip_is_byte_exact(duck_pt, tmp_path):
    """reference tests/pt_format.cpp:18-178: serialize -> deserialize round-trips every array and texture."""
    raw = O.duck_pt_bytes()
    assert raw[:9] == b"PTFORMAT3"
    assert duck_pt.dumps() == raw
    path = tmp_path / "duck.pt"
    duck_pt.save(path)
    again = rf.PtFormat.load(path)
    for name in rf.PtFormat.ARRAYS:
        assert getattr(again, name).tobytes() == getattr(duck_pt, name).tobytes(), name
    assert len(again.base_color_textures) == 1
    assert again.base_color_textures[0].tobytes() == duck_pt.base_color_textures[0].tobytes()
    # layout: magic, then [u64 n][n x 48 B] nodes
    n = int(np.frombuffer(raw[9:17], dtype="<u8")[0])
    assert n == duck_pt.bvh_nodes.size and raw[17:17 + 48] == duck_pt.bvh_nodes[:1].tobytes()


def test_pt_empty_round_trip():
    empty = rf.PtFormat()
    raw = empty.dumps()
    assert len(raw) == 9 + 13 * 8 + 8
    again = rf.PtFormat.loads(raw)
    assert again.bvh_nodes.size == 0 and again.base_color_textures == []


def test_pt_invalid_magic_messages():
    """reference tests/pt_format.cpp:180-213 — exact error strings."""
    with pytest.raises(rf.RayfinderError) as e:
        rf.PtFormat.loads(b"PTFORMAT0")
    assert str(e.value) == ("Mismatching PtFormat file version. Invalid version in magic bytes: expected "
                            "'PTFORMAT3', got 'PTFORMAT0'.")
    assert e.value.status == capi.RF_ERROR_FORMAT
    with pytest.raises(rf.RayfinderError) as e:
        rf.PtFormat.loads(b"INVALID  ")
    assert str(e.value) == "Invalid file format: expected PtFormat file."


def test_pt_truncated_and_missing_file(tmp_path):
    raw = O.duck_pt_bytes()
    for cut in (4, 9, 12, 17 + 48 * 10 + 5, len(raw) - 1):
        with pytest.raises(rf.RayfinderError):
            rf.PtFormat.loads(raw[:cut])
    with pytest.raises(rf.RayfinderError) as e:
        rf.PtFormat.load(tmp_path / "nope.pt")
    assert str(e.value) == f"Failed to open file: {tmp_path / 'nope.pt'}"  # common/file_stream.cpp:13-16


def test_pt_texture_pixel_count_must_match_its_size():
    """A texture whose numPixels differs from width x height would make every consumer (which reads width x height
    texels) run past the pixel array: rejected as malformed; a count whose byte size wraps 64 bits is a short read."""
    raw = bytearray(O.duck_pt_bytes())
    pt = rf.PtFormat.loads(bytes(raw))
    tex = pt.base_color_textures[0]
    header = len(raw) - tex.size * 4 - 16  # [u32 width][u32 height][u64 numPixels] of the only texture
    assert np.frombuffer(raw[header:header + 8], dtype="<u4").tolist() == [tex.shape[1], tex.shape[0]]
    grown = bytearray(raw)
    grown[header:header + 4] = np.uint32(tex.shape[1] * 2).tobytes()  # width doubled, pixel count unchanged
    with pytest.raises(rf.RayfinderError) as e:
        rf.PtFormat.loads(bytes(grown))
    assert e.value.status == capi.RF_ERROR_FORMAT and "pixels for" in str(e.value)
    wrapped = bytearray(raw)
    wrapped[header + 8:header + 16] = np.uint64(1 << 62).tobytes()  # n * 4 == 0 (mod 2^64)
    with pytest.raises(rf.RayfinderError) as e:
        rf.PtFormat.loads(bytes(wrapped))
    assert e.value.status == capi.RF_ERROR_IO


def test_reorder_attributes():
    idx = np.array([2, 0, 1], dtype=np.uint64)
    assert rf.reorder_attributes(np.array([10, 20, 30]), idx).tolist() == [20, 30, 10]  # bvh.hpp:36-46

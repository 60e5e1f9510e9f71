"""GPU buildBvh (csrc/bvh_build.cu) against the host builder rf_build_bvh, which
tests/test_oracle_vs_reference.py pins byte for byte to the reference's own common/bvh.cpp.

Bar: byte-identical BvhNode arrays (boxes with their signed zeros, offsets, counts, split axes, padding) and
identical triangleIndices, for real scenes and for soups built to hit every branch of bvh.cpp:96-221."""
import numpy as np
import pytest

import _oracle as O
import rayfinder_b200 as rf

pytestmark = pytest.mark.gpu


def assert_same_bvh(tris):
    nodes_h, idx_h = rf.build_bvh(tris)
    nodes_d, idx_d, ms = rf.build_bvh_device(tris)
    assert nodes_d.shape == nodes_h.shape
    assert nodes_d.tobytes() == nodes_h.tobytes()
    assert np.array_equal(idx_d, idx_h)
    return nodes_h, ms


def test_duck_and_sponza_byte_identical(duck_pt, sponza_pt):
    for pt, leaves_max in ((duck_pt, 255), (sponza_pt, 255)):
        tris = O.triangles9(pt)
        # the .pt stores the triangles in leaf order; shuffle them so the build starts from an arbitrary soup
        perm = np.random.default_rng(5).permutation(len(tris))
        nodes, ms = assert_same_bvh(tris[perm])
        assert nodes["triangle_count"].max() <= leaves_max and ms > 0.0
    # and in file order (the order the baker's flattened model would hand over)
    assert_same_bvh(O.triangles9(duck_pt))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 33, 257, 1000, 40000])
def test_random_soups(n):
    rng = np.random.default_rng(n)
    centres = rng.uniform(-10, 10, size=(n, 1, 3))
    tris = (centres + rng.normal(scale=0.3, size=(n, 3, 3))).astype(np.float32)
    assert_same_bvh(tris)


def test_ties_signed_zeros_and_flat_boxes():
    """Coordinates on a coarse grid (equal centroids, equal bucket boundaries, zero-area nodes) with -0.0f and +0.0f
    mixed in: the sequential `(b < a) ? b : a` folds of the reference decide which zero a box keeps."""
    rng = np.random.default_rng(11)
    for n, grid in ((300, 2), (5000, 3), (20000, 8)):
        tris = rng.integers(-grid, grid + 1, size=(n, 3, 3)).astype(np.float32)
        neg = rng.random(size=tris.shape) < 0.5
        tris = np.where((tris == 0) & neg, np.float32(-0.0), tris).astype(np.float32)
        flat = rng.random(n) < 0.3
        tris[flat, :, 1] = np.where(rng.random((flat.sum(), 3)) < 0.5, np.float32(-0.0), np.float32(0.0))  # flat at y = +-0
        assert_same_bvh(tris)


def test_forced_splits_and_big_leaves():
    """More than 255 primitives with identical centroids stay one leaf (cLo == cHi, bvh.cpp:111); more than 255 whose
    SAH says 'leaf' are split anyway (bvh.cpp:206-207)."""
    rng = np.random.default_rng(3)
    same = np.tile(rng.normal(size=(1, 3, 3)), (700, 1, 1)).astype(np.float32)
    nodes, _ = assert_same_bvh(same)
    assert len(nodes) == 1 and nodes["triangle_count"][0] == 700
    # big overlapping triangles, centroids spread a little: the SAH prefers a leaf, the 255 limit forces splits
    big = (rng.normal(scale=100.0, size=(3000, 3, 3)) + rng.normal(scale=0.01, size=(3000, 1, 3))).astype(np.float32)
    nodes, _ = assert_same_bvh(big)
    assert nodes["triangle_count"].max() <= 255
    assert_same_bvh(np.concatenate([same, big]))


def test_device_built_bvh_traces_like_the_host_built_one(duck_pt):
    tris = O.triangles9(duck_pt)
    nodes, idx, _ = rf.build_bvh_device(tris)
    ordered = rf.reorder_attributes(tris, idx)
    scene = rf.TraversalScene(nodes, ordered)
    cam = rf.bvh_visualizer_camera(nodes, 256, 256)
    counts, _ = scene.bvh_visualizer_node_counts(cam, 256, 256)
    expected, _ = O.oracle_node_counts(nodes, ordered, rf.camera_to_array(cam), 256, 256, rf.FLT_MAX)
    assert np.array_equal(counts, expected)

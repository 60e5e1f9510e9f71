"""The C-ABI library loads, exports every symbol include/rayfinder_b200.h declares, and fails loudly —
never falls back — when no CUDA device is present.  No compute calls here."""
import ctypes as C
import re
import subprocess

import numpy as np
import pytest

import _oracle as O
import rayfinder_b200 as rf
from rayfinder_b200 import capi

HEADER = O.ROOT / "include" / "rayfinder_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 35
    assert sorted(capi.SIGNATURES) == names  # the ctypes table covers the header exactly
    lib = capi.lib()
    for name in names:
        assert getattr(lib, name) is not None
    exported = subprocess.run(["nm", "-D", "--defined-only", str(capi.LIB_PATH)], capture_output=True, text=True).stdout
    for name in names:
        assert re.search(rf"\bT {name}\b", exported), name


def test_library_carries_sm100a_kernels():
    assert capi.lib().rf_has_cuda_kernels() == 1
    assert b"sm_100a" in capi.lib().rf_build_info()
    out = subprocess.run(["cuobjdump", "-lelf", str(capi.LIB_PATH)], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout


def test_oracle_is_not_linked_into_the_product():
    needed = subprocess.run(["ldd", str(capi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "oracle" not in needed
    sources = list((O.ROOT / "rayfinder_b200").glob("*.py")) + list((O.ROOT / "rayfinder_b200" / "csrc").glob("*"))
    assert len(sources) > 8
    for src in sources:
        text = src.read_text()
        for needle in ("liboracle", "libref_oracle", "_oracle", "oracle/", "oracle."):
            assert needle not in text, (src, needle)


def _cuda_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_cuda_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device(duck_pt):
    with pytest.raises(rf.RayfinderError) as e:
        rf.TraversalScene(duck_pt.bvh_nodes, O.triangles9(duck_pt))
    assert e.value.status == capi.RF_ERROR_CUDA and "no CPU fallback" in str(e.value)
    params = rf.RenderParameters((64, 64), rf.bvh_visualizer_camera(duck_pt.bvh_nodes, 64, 64), rf.SamplingParams(1, 2))
    with pytest.raises(rf.RayfinderError) as e:
        rf.ReferencePathTracer(params, (64, 64), rf.SceneArrays.from_pt(duck_pt))
    assert e.value.status == capi.RF_ERROR_CUDA
    with pytest.raises(rf.RayfinderError) as e:
        rf.build_bvh_device(O.triangles9(duck_pt)[:100])
    assert e.value.status == capi.RF_ERROR_CUDA and "no CPU fallback" in str(e.value)


def test_renderer_argument_validation(duck_pt):
    """Errors that are detected before any device work."""
    scene = rf.SceneArrays.from_pt(duck_pt)
    cam = rf.bvh_visualizer_camera(duck_pt.bvh_nodes, 64, 64)
    with pytest.raises(rf.RayfinderError) as e:  # framebuffer larger than maxFramebufferSize
        rf.ReferencePathTracer(rf.RenderParameters((128, 64), cam, rf.SamplingParams(1, 2)), (64, 64), scene)
    assert e.value.status == capi.RF_ERROR_INVALID_ARGUMENT
    bad = duck_pt.bvh_nodes.copy()
    bad["second_child_offset"][0] = 0  # would loop forever on the device
    with pytest.raises(rf.RayfinderError) as e:
        rf.TraversalScene(bad, O.triangles9(duck_pt))
    assert "child indices" in str(e.value)
    bad = duck_pt.bvh_nodes.copy()
    leaf = int(np.flatnonzero(bad["triangle_count"] > 0)[0])
    bad["triangles_offset"][leaf] = 10_000_000
    with pytest.raises(rf.RayfinderError) as e:
        rf.TraversalScene(bad, O.triangles9(duck_pt))
    assert "out of range" in str(e.value)
    # a DAG (every interior node i points at i + 1 and i + 2): structurally "ordered", but a node with two parents would
    # make the depth walk enumerate Fibonacci-many paths — it must be rejected at once
    import time
    n = 60  # (the longest descent stays below the 32-entry stack limit, so that check does not fire first)
    dag = np.zeros(n, dtype=rf.BVH_NODE_DTYPE)
    dag["aabb_max"] = 1.0
    dag["second_child_offset"] = np.arange(n) + 2
    dag["triangle_count"][n - 2:] = 1
    dag["split_axis"][n - 2:] = 0xFFFFFFFF
    t0 = time.perf_counter()
    with pytest.raises(rf.RayfinderError) as e:
        rf.TraversalScene(dag, O.triangles9(duck_pt))
    assert "more than one parent" in str(e.value) and time.perf_counter() - t0 < 1.0

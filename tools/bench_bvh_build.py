#!/usr/bin/env python3
"""buildBvh: host builder (rf_build_bvh) against the GPU builder (rf_build_bvh_device) on the baked scenes (GPU box)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import _oracle as O  # noqa: E402
import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402

for name in ("Duck", "Sponza"):
    pt = rf.PtFormat.loads(O.duck_pt_bytes()) if name == "Duck" else rfa.load_scene(name)
    tris = O.triangles9(pt)
    tris = tris[np.random.default_rng(1).permutation(len(tris))]
    t0 = time.perf_counter()
    nodes_h, idx_h = rf.build_bvh(tris)
    host_s = time.perf_counter() - t0
    from rayfinder_b200 import capi

    for mode, label in ((0, "one persistent launch"), (1, "one launch per phase and level")):
        capi.lib().rf_build_bvh_device_set_mode(mode)
        rf.build_bvh_device(tris)  # warm-up (context, allocations)
        best_ms, best_wall = 1e9, 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            nodes_d, idx_d, ms = rf.build_bvh_device(tris)
            best_wall = min(best_wall, time.perf_counter() - t0)
            best_ms = min(best_ms, ms)
        same = nodes_d.tobytes() == nodes_h.tobytes() and np.array_equal(idx_d, idx_h)
        print(f"{name}: {len(tris)} triangles -> {len(nodes_h)} nodes; host {host_s * 1e3:.1f} ms (1 thread); device, {label}: {best_ms:.2f} ms "
              f"(kernels), {best_wall * 1e3:.1f} ms (call incl. copies and allocation); byte-identical: {same}; "
              f"{len(tris) / best_ms / 1e3:.1f} M triangles/s", flush=True)
        if mode == 0:
            import ctypes as C

            phases = (C.c_float * 12)()
            levels = capi.lib().rf_build_bvh_device_last_phases(phases)
            names = ("boxes", "decide", "buckets", "sweep", "scan", "-", "pair", "permute", "-", "leaf scan", "emit", "block-local subtrees")
            print(f"    {levels} grid-wide levels; ms per phase (block 0, incl. the grid barrier): " + ", ".join(f"{k} {v:.3f}" for k, v in zip(names, phases)), flush=True)
    capi.lib().rf_build_bvh_device_set_mode(0)

#!/usr/bin/env python3
"""Latency of single rays through the persistent traversal loop (GPU box, instrumented debug build).

    python -m rayfinder_b200._build --timeline
    python tools/lone_ray.py

Traces grazing rays of the Sponza interior (the kind that makes the tail of a traversal launch), finds the
longest ones, then traces each ALONE (1 ray in the whole launch), 32 copies in one warp, and 32 different long rays in
one warp, and reports ns per (node visit + triangle test) from the per-warp %globaltimer records.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("RAYFINDER_B200_LIB", str(ROOT / "rayfinder_b200" / "librayfinder_b200_timeline.so"))

import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402
from rayfinder_b200 import capi  # noqa: E402

REC = np.dtype([("tag", "<u8"), ("start", "<u8"), ("dry", "<u8"), ("exit", "<u8"), ("rays", "<u4"), ("rounds", "<u4"), ("sm", "<u4"), ("pad", "<u4")])


def main():
    capi.lib()
    raw = C.CDLL(str(capi.LIB_PATH))
    raw.rf_debug_timeline_arm.argtypes = [C.c_uint32]
    raw.rf_debug_timeline_read.argtypes = [C.c_void_p, C.c_uint32]
    raw.rf_debug_timeline_read.restype = C.c_uint32
    pt = rfa.load_scene("Sponza")
    scene = rf.TraversalScene(pt.bvh_nodes, pt.bvh_position_attributes)
    rng = np.random.default_rng(7)
    n = 200000
    # origins inside the atrium, directions mostly horizontal (grazing along floors and walls)
    lo, hi = pt.bvh_nodes["aabb_min"][0], pt.bvh_nodes["aabb_max"][0]
    o = lo + (hi - lo) * rng.uniform(0.2, 0.8, size=(n, 3))
    d = rng.normal(size=(n, 3)) * np.array([1.0, 0.05, 1.0])
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    hit, p_t, nodes = scene.ray_intersect_bvh(rays, 10000.0)
    order = np.argsort(nodes)[::-1]
    print(f"{n} rays: mean nodes {nodes.mean():.1f}, p99 {np.percentile(nodes, 99):.0f}, max {nodes.max()}")

    def timed(batch, label):
        best = None
        for _ in range(3):
            scene.ray_intersect_bvh(rays, 10000.0)  # keep the clocks boosted and the scene L2-resident
            assert raw.rf_debug_timeline_arm(1 << 16) == 0
            _, _, nv = scene.ray_intersect_bvh(batch, 10000.0)
            buf = np.zeros(1 << 16, dtype=REC)
            k = raw.rf_debug_timeline_read(buf.ctypes.data, 1 << 16)
            r = buf[:k]
            r = r[r["rays"] > 0]
            span = (r["exit"].astype(np.int64) - r["start"].astype(np.int64)).max()
            best = span if best is None else min(best, span)
        tris = int(r["pad"].max() >> 16)
        print(f"{label:42s}: {best / 1e3:8.1f} us, longest ray {nv.max():5d} nodes + {tris:4d} tris -> {best / (nv.max() + tris):6.1f} ns per op "
              f"({best / (nv.max() + tris) * 1.965:5.0f} cycles @1965 MHz)", flush=True)

    for k in range(3):
        timed(rays[order[k]][None, :], f"lone ray #{k}")
    timed(np.repeat(rays[order[0]][None, :], 32, axis=0), "32 copies of ray #0 in one warp")
    timed(rays[order[:32]], "the 32 longest rays in one warp")
    timed(rays[order[:32 * 148:148]], "32 long rays (every 148th) in one warp")
    timed(rays[order[:4736]], "4736 longest rays (148 warps x 32)")


if __name__ == "__main__":
    main()

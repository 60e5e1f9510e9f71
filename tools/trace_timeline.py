#!/usr/bin/env python3
"""Per-warp timeline of the traversal launches of one frame (GPU box, instrumented debug build).

    python -m rayfinder_b200._build --timeline
    RAYFINDER_B200_LIB=rayfinder_b200/librayfinder_b200_timeline.so python tools/trace_timeline.py [WxH] [sub_frames] [evict_max] [persistent_kernel]

For every traversal launch: when the first / last warp started, when the ray queue ran dry (first / last warp to
notice), when the warps exited (percentiles), and how many warps were still running at points of the tail.
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
    size = sys.argv[1] if len(sys.argv) > 1 else "672x384"
    sub = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    evict = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    mega = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    w, h = (int(x) for x in size.split("x"))
    lib = capi.lib()
    raw = C.CDLL(str(capi.LIB_PATH))
    raw.rf_debug_timeline_arm.argtypes = [C.c_uint32]
    raw.rf_debug_timeline_read.argtypes = [C.c_void_p, C.c_uint32]
    raw.rf_debug_timeline_read.restype = C.c_uint32
    pt = rfa.load_scene("Sponza")
    params = rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, 8), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (w, h), rf.SceneArrays.from_pt(pt))
    ren.set_pipeline(sub, mega, 3, 256)
    ren.set_tail_policy(evict)
    for k in range(3):
        params.exposure = 0.25 + 0.01 * k
        ren.set_render_parameters(params)
        ren.render()
    ren.synchronize()
    cap = 1 << 20
    assert raw.rf_debug_timeline_arm(cap) == 0
    params.exposure = 0.3
    ren.set_render_parameters(params)
    ren.render()
    ren.synchronize()
    buf = np.zeros(cap, dtype=REC)
    n = raw.rf_debug_timeline_read(buf.ctypes.data, cap)
    rec = buf[:n]
    per_ray = rec[rec["tag"] == 1]
    rec = rec[rec["tag"] != 1]
    if len(per_ray):
        dur = (per_ray["exit"].astype(np.int64) - per_ray["start"].astype(np.int64)) / 1e3
        ops = per_ray["rays"].astype(np.int64) + per_ray["pad"]
        print(f"straggler rays: {len(per_ray)}; nodes/ray mean {per_ray['rays'].mean():.1f} max {per_ray['rays'].max()}; tris/ray mean {per_ray['pad'].mean():.1f}; "
              f"nodes per window {per_ray['rays'].sum() / max(1, per_ray['rounds'].sum()):.2f}; us/ray mean {dur.mean():.2f} max {dur.max():.1f}; "
              f"ns per op overall {dur.sum() * 1e3 / max(1, ops.sum()):.1f}")
        for j in np.argsort(dur)[-5:]:
            print(f"   long ray: {dur[j]:.1f} us, {int(per_ray['rays'][j])} nodes, {int(per_ray['pad'][j])} tris, {int(per_ray['rounds'][j])} windows -> {dur[j] * 1e3 / max(1, ops[j]):.0f} ns/op")
    t0 = int(rec["start"].min())
    print(f"{w}x{h} sub_frames={sub} evict_max={evict}: {n} warp records, frame span {(int(rec['exit'].max()) - t0) / 1e3:.1f} us")
    order = sorted(set(rec["tag"].tolist()), key=lambda t: int(rec["start"][rec["tag"] == t].min()))
    print("launch  warps     rays | start(first..last)  dry(first..last)  exit p50   p90   p99   max  | warps alive after dry+0/50/100/200/300us | rounds max")
    for i, tag in enumerate(order):
        r = rec[rec["tag"] == tag]
        s0 = int(r["start"].min())
        rel = lambda a: (a.astype(np.int64) - s0) / 1e3  # noqa: E731
        start, ex = rel(r["start"]), rel(r["exit"])
        dry = rel(r["dry"][r["dry"] != 0]) if (r["dry"] != 0).any() else np.array([0.0])
        alive = [int((ex > dry.min() + d).sum()) for d in (0, 50, 100, 200, 300)]
        print(f"{i:3d} @{(s0 - t0) / 1e3:8.1f} {len(r):5d} {int(r['rays'].sum()):8d} | {start.min():6.1f} {start.max():6.1f}   {dry.min():7.1f} {dry.max():7.1f}   "
              f"{np.percentile(ex, 50):7.1f} {np.percentile(ex, 90):7.1f} {np.percentile(ex, 99):7.1f} {ex.max():7.1f} | {alive} | {int(r['rounds'].max())}")
        last = np.argsort(r["exit"])[-4:]
        print("        last warps: " + "; ".join(f"exit {ex[j]:.0f}us rays {int(r['rays'][j])} rounds {int(r['rounds'][j])} maxnodes {int(r['pad'][j]) & 0xFFFF} maxtris {int(r['pad'][j]) >> 16} sm {int(r['sm'][j])}" for j in last))
    ren.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Deferred lighting pass at 1920x1080 on Sponza (GPU box): ms per frame through the C-ABI with the host G-buffer copied in
every frame, and the rays it traces.  The G-buffer is built from the primary hits of the interior view (as in the tests)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402
from test_gpu_deferred import make_gbuffer  # noqa: E402

w, h = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1920x1080").split("x"))
pt = rfa.load_scene("Sponza")
eye, centre = np.array([1.22, 1.25, -1.25]), np.array([-5.0, 0.5, 6.0])
inv, albedo, normal, depth = make_gbuffer(pt, w, h, eye, centre, seed=7)
ren = rf.ReferencePathTracer(rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, 2)), (w, h), rf.SceneArrays.from_pt(pt))
for k in range(3):
    ren.render_deferred_lighting(inv, eye, k, albedo, normal, depth)
ren.synchronize()
ren.reset_stats()
frames = 20
t0 = time.perf_counter()
for k in range(frames):
    ren.render_deferred_lighting(inv, eye, 3 + k, albedo, normal, depth)
ren.synchronize()
dt = time.perf_counter() - t0
s = ren.stats()
rays = s["closest_rays"] + s["shadow_rays"]
print(f"{w}x{h}: {s['device_ms_total'] / frames:.2f} ms per frame on the device ({rays / s['device_ms_total'] / 1e3:.0f} Mrays/s), "
      f"{dt / frames * 1e3:.2f} ms per frame end to end ({(albedo.nbytes + normal.nbytes + depth.nbytes) / 1e6:.0f} MB of G-buffer copied in per frame), "
      f"{rays // frames} rays per frame ({s['shadow_rays'] // frames} shadow, {s['closest_rays'] // frames} closest) -> {rays / dt / 1e6:.0f} Mrays/s end to end")

#!/usr/bin/env python3
"""Per-launch spans of one staged frame (one tile set, one stream) at a few frame sizes (GPU box).

    python tools/stage_times.py [WxH ...]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402

pt = rfa.load_scene("Sponza")
scene = rf.SceneArrays.from_pt(pt)
for size in (sys.argv[1:] or ["1920x1080", "672x384"]):
    w, h = (int(x) for x in size.split("x"))
    params = rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, 8), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (w, h), scene)
    ren.set_pipeline(1, 0, 3, 256)
    ren.set_stage_timing(True)
    ren.set_option("stage_debug", 1)
    print(f"== {w}x{h}", file=sys.stderr, flush=True)
    for k in range(4):
        params.exposure = 0.25 + 0.01 * k
        ren.set_render_parameters(params)
        ren.render()
        ren.synchronize()
    s = ren.stats()
    print({k: s[k] for k in ("closest_rays", "shadow_rays", "closest_nodes_visited", "shadow_nodes_visited")}, file=sys.stderr)
    ren.close()

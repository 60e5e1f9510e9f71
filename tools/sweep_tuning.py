#!/usr/bin/env python3
"""Sweep the scheduling knobs of the persistent traversal kernels on one renderer (GPU box).
Prints one line per setting: per-stage milliseconds of a Sponza 1080p / 8-bounce frame."""
import itertools
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402


def main():
    w, h, bounces, frames = 1920, 1080, 8, 4
    pt = rfa.load_scene("Sponza")
    params = rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, bounces), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (w, h), rf.SceneArrays.from_pt(pt))
    ren.set_stage_timing(True)
    tri_list = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1,4,8,12,16,24,32".split(","))]
    refill_list = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,4,8,16,32".split(","))]
    blocks_list = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "4".split(","))]
    for blocks, tri, refill in itertools.product(blocks_list, tri_list, refill_list):
        ren.set_tuning(tri, refill, blocks)
        for k in range(frames + 1):
            if k == 1:
                ren.reset_stats()
            params.exposure = 0.25 + 0.01 * k
            ren.set_render_parameters(params)
            ren.render()
        s = ren.stats()
        rays = s["closest_rays"] + s["shadow_rays"]
        print(f"blocks={blocks} tri_min={tri:2d} refill_min={refill:2d}  total={s['device_ms_total'] / frames:7.3f} ms  "
              f"closest={s['device_ms_closest'] / frames:7.3f} shadow={s['device_ms_shadow'] / frames:7.3f} "
              f"shade={s['device_ms_shade'] / frames:6.3f}  Mrays/s={rays / s['device_ms_total'] / 1e3:8.1f}", flush=True)


if __name__ == "__main__":
    main()

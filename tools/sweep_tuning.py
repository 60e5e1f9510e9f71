#!/usr/bin/env python3
"""Sweep the scheduling knobs of the persistent traversal kernels on one renderer (GPU box).
Prints one line per setting: per-stage milliseconds of a Sponza 1080p / 8-bounce frame.

    python tools/sweep_tuning.py TRI_LIST REFILL_LIST BLOCKS_LIST VARIANT_LIST SUBFRAMES_MINUS_1_LIST      (comma-separated)
"""
import itertools
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402


def ints(idx, default):
    return [int(x) for x in (sys.argv[idx] if len(sys.argv) > idx else default).split(",")]


def main():
    import os
    w, h, bounces, frames = int(os.environ.get('RF_W', 1920)), int(os.environ.get('RF_H', 1080)), 8, 4
    pt = rfa.load_scene("Sponza")
    params = rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, bounces), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (w, h), rf.SceneArrays.from_pt(pt))
    ren.set_stage_timing(os.environ.get('RF_STAGE_TIMING', '0') == '1')
    tri_list, refill_list = ints(1, "2,4,8"), ints(2, "4")
    blocks_list, variant_list, sort_list, bs_list, mega_list = ints(3, "4"), ints(4, "3"), ints(5, "1"), ints(6, "256"), ints(7, "1")
    for mega, bs, sort, variant, blocks, tri, refill in itertools.product(mega_list, bs_list, sort_list, variant_list, blocks_list, tri_list, refill_list):
        ren.set_tuning(tri, refill, blocks)
        ren.set_pipeline(sort + 1, mega, variant, bs)
        ren.set_tail_policy(int(os.environ.get('RF_EVICT_MAX', -1)))
        for k in range(frames + 1):
            if k == 1:
                ren.reset_stats()
            params.exposure = 0.25 + 0.01 * k
            ren.set_render_parameters(params)
            ren.render()
        s = ren.stats()
        rays = s["closest_rays"] + s["shadow_rays"]
        print(f"mega={mega} block={bs} subframes={sort + 1} variant={variant:2d} (steps={(variant & 3) + 1} leaf={(variant >> 2) & 1} branchy={(variant >> 3) & 1}) "
              f"blocks={blocks} tri_min={tri:2d} refill_min={refill:2d}  total={s['device_ms_total'] / frames:7.3f} ms  "
              f"trace={s['device_ms_trace'] / frames:7.3f} "
              f"shade={s['device_ms_shade'] / frames:6.3f}  Mrays/s={rays / s['device_ms_total'] / 1e3:8.1f}", flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""A/B of the traversal kernels and their scheduling knobs on one GPU (Sponza, 8 bounces): ms per frame and Mrays/s.

    python tools/kernel_ab.py [WxH ...] [--configs "kernel:pair_variant:sub_frames:blocks:tri_min:refill_min[:evict_max[:persistent]],..."] [--frames N]

kernel 1 = one node per visit (traversal.cuh), 2 = child-pair records (traversal_pairs.cuh); sub_frames -1 = automatic.
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402

DEFAULT = "1:3:-1:4:4:4,2:3:-1:4:4:4,2:7:-1:4:4:4,2:1:-1:4:4:4,2:5:-1:4:4:4,2:3:1:4:4:4,2:7:1:4:4:4,2:7:2:4:4:4,2:7:-1:4:2:4,2:7:-1:4:8:4,2:7:-1:4:4:8,2:7:-1:4:4:2"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sizes", nargs="*", default=["1920x1080", "672x384"])
    ap.add_argument("--configs", default=DEFAULT)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--options", default="", help="extra rf_renderer_set_option settings for every config: name=value,...")
    args = ap.parse_args()
    pt = rfa.load_scene("Sponza")
    scene = rf.SceneArrays.from_pt(pt)
    for size in args.sizes:
        w, h = (int(x) for x in size.split("x"))
        params = rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, 8), rf.Sky(), 0.25)
        ren = rf.ReferencePathTracer(params, (w, h), scene)
        for cfg in args.configs.split(","):
            fields = [int(x) for x in cfg.split(":")]
            kernel, variant, sub_frames, blocks, tri_min, refill_min = fields[:6]
            evict = fields[6] if len(fields) > 6 else -1
            persistent = fields[7] if len(fields) > 7 else 0
            ren.set_option("trace_kernel", kernel)
            if kernel == 2:
                ren.set_option("pair_variant", variant)
                ren.set_pipeline(sub_frames, persistent, -1, 0)
            else:
                ren.set_pipeline(sub_frames, persistent, variant, 256)
            for item in filter(None, args.options.split(",")):
                name, value = item.split("=")
                ren.set_option(name, int(value))
            ren.set_tuning(tri_min, refill_min, blocks)
            ren.set_tail_policy(evict)
            for k in range(args.frames + 2):
                if k == 2:
                    ren.reset_stats()
                params.exposure = 0.25 + 0.01 * (k % 7)
                ren.set_render_parameters(params)
                ren.render()
            s = ren.stats()
            rays = s["closest_rays"] + s["shadow_rays"]
            print(f"{w}x{h} persistent={persistent} kernel={kernel} variant={variant} sub_frames={s['sub_frames']} evict={s['evict_max']} blocks={blocks} tri_min={tri_min} "
                  f"refill_min={refill_min}: {s['device_ms_total'] / args.frames:7.3f} ms/frame  {rays / s['device_ms_total'] / 1e3:8.1f} Mrays/s  "
                  f"records/ray {s['node_records_loaded'] / rays:6.2f}  visits/ray {(s['closest_nodes_visited'] + s['shadow_nodes_visited']) / rays:6.2f}", flush=True)
        ren.close()


if __name__ == "__main__":
    main()

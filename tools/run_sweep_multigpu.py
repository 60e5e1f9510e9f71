#!/usr/bin/env python3
"""BASELINE.json configs[4], the multi-GPU part: ray-throughput sweep 1/2/4/8/16 bounces x 256^2 .. 4096^2 on N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/run_sweep_multigpu.py [--quick] [--out gpurun_out/sweep_8gpu.json]

(also runs as plain `python tools/run_sweep_multigpu.py` = one GPU, the same loop without the exchange partner.)  Every rank
renders its tiles with the automatic schedule, the frame is exchanged to rank 0 inside the timed region, times are CUDA events
on the rendering stream, max over ranks; rank 0 writes one JSON table.  Not a bench value — bench.py is; this is the sweep
BASELINE.json asks for.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402
from rayfinder_b200 import distributed as rfd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sizes = (256, 1024) if args.quick else (256, 512, 1024, 2048, 4096)
    bounce_counts = (1, 8) if args.quick else (1, 2, 4, 8, 16)
    top = max(sizes)
    pt = rfa.load_scene("Sponza")
    params = rf.RenderParameters((top, top), rf.fly_camera(top, top), rf.SamplingParams(1, 8), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (top, top), rf.SceneArrays.from_pt(pt), device=local)
    stream = torch.cuda.current_stream(dev)
    ren.set_stream(stream.cuda_stream)
    ren.set_tile_partition(rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    table = []
    for size in sizes:
        exchange = rfd.HdrExchange(ren, size, size, mode="auto")
        for bounces in bounce_counts:
            params = rf.RenderParameters((size, size), rf.fly_camera(size, size), rf.SamplingParams(1, bounces), rf.Sky(), 0.25)
            starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.frames)]
            ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.frames)]
            for k in range(args.frames + 2):
                if k == 2:
                    ren.synchronize()
                    ren.reset_stats()
                params.exposure = 0.25 + 0.01 * k  # restarts the accumulation: every step traces the frame again
                ren.set_render_parameters(params)
                flush.zero_()
                if k >= 2:
                    starts[k - 2].record(stream)
                ren.render()
                exchange()
                if k >= 2:
                    ends[k - 2].record(stream)
            torch.cuda.synchronize(dev)
            s = ren.stats()
            agg = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(starts, ends))], dtype=torch.float64, device=dev)
            cnt = torch.tensor([s["closest_rays"] + s["shadow_rays"], s["paths"]], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(agg, op=dist.ReduceOp.MAX)
                dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            ms = float(agg[0]) / args.frames
            rays = int(cnt[0]) / args.frames
            row = {"size": size, "bounces": bounces, "n_gpus": world, "ms_per_frame": ms, "rays_per_frame": rays, "mrays_s": rays / ms / 1e3,
                   "paths_per_frame": int(cnt[1]) / args.frames, "persistent_kernel": s["persistent_kernel"], "sub_frames": s["sub_frames"],
                   "evict_max": s["evict_max"], "exchange": exchange.mode}
            table.append(row)
            if rank == 0:
                print("sweep", row, flush=True)
        if world > 1:
            dist.barrier()
        exchange.close()
    if rank == 0:
        out = Path(args.out) if args.out else ROOT / "gpurun_out" / f"sweep_{world}gpu.json"
        out.parent.mkdir(exist_ok=True)
        out.write_text(json.dumps({"scene": "Sponza.pt", "gpu": "B200", "n_gpus": world, "frames_per_point": args.frames,
                                   "timing": "CUDA events around render + exchange, L2 flushed before every frame, max over ranks", "sweep": table}, indent=1) + "\n")
    ren.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

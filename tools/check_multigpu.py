#!/usr/bin/env python3
"""Multi-GPU correctness check (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_multigpu.py [width height bounces]

Every rank renders its tiles of a Sponza frame, the HDR buffer is brought together on rank 0 with each exchange mode
(NCCL sum-reduce; stores over NVLink peer memory fused into the accumulation kernel), and rank 0 compares the result bit
for bit with a single-GPU render of the whole frame (BASELINE.json
configs[3] at 3840x2160 by default)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402
from rayfinder_b200 import distributed as rfd  # noqa: E402


def main():
    w, h, bounces = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (3840, 2160, 8)))
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pt = rfa.load_scene("Sponza")
    params = rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, bounces), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (w, h), rf.SceneArrays.from_pt(pt), device=local)
    ren.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    ren.set_tile_partition(rank, world)
    spp = 3  # progressive: the exchange runs after every frame and must hand over the k-sample sum each time
    params.sampling_params = rf.SamplingParams(spp, bounces)
    results = {}
    for mode in ("nccl", "p2p"):
        exchange = rfd.HdrExchange(ren, w, h, mode=mode)
        params.exposure = 0.25 if mode == "nccl" else 0.35  # restarts the accumulation while the buffers are in use
        ren.set_render_parameters(params)
        ren.set_frame_count(0)
        ren.reset_stats()
        frames = []
        for frame in range(spp):
            ren.render()
            hdr = exchange()
            if rank == 0:
                frames.append(hdr.cpu().numpy().copy())
        torch.cuda.synchronize(dev)
        paths = torch.tensor([ren.stats()["paths"]], dtype=torch.int64, device=dev)
        dist.all_reduce(paths)
        if rank == 0:
            results[exchange.mode] = (frames, int(paths[0]))
        dist.barrier()
        exchange.close()
    if rank == 0:
        ren.set_tile_partition(0, 1)
        params.exposure = 0.9
        ren.set_render_parameters(params)
        ren.set_frame_count(0)
        singles = []
        for frame in range(spp):
            ren.render()
            singles.append(ren.read_hdr()[0])
        for mode, (frames, paths) in results.items():
            same = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(frames, singles))
            print(f"world={world} {w}x{h} bounces={bounces} spp={spp} exchange={mode}: paths={paths} (expected {spp * w * h}), "
                  f"image on rank 0 after every frame bit-identical to single GPU: {same}", flush=True)
            assert same and paths == spp * w * h
        assert "p2p" in results or world == 1, "peer-memory exchange was not available"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

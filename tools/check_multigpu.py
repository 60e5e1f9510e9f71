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
    results = {}
    for mode in ("nccl", "p2p"):
        exchange = rfd.HdrExchange(ren, w, h, mode=mode)
        for frame in range(2):  # two frames: the second restarts the accumulation while the buffers are in use
            params.exposure = 0.25 + 0.1 * frame
            ren.set_render_parameters(params)
            ren.reset_stats()
            ren.render()
            hdr = exchange()
        torch.cuda.synchronize(dev)
        paths = torch.tensor([ren.stats()["paths"]], dtype=torch.int64, device=dev)
        dist.all_reduce(paths)
        if rank == 0:
            results[exchange.mode] = (hdr.cpu().numpy().copy(), int(paths[0]))
        dist.barrier()
        exchange.close()
    if rank == 0:
        ren.set_tile_partition(0, 1)
        params.exposure = 0.9
        ren.set_render_parameters(params)
        ren.render()
        single, _ = ren.read_hdr()
        for mode, (image, paths) in results.items():
            same = np.array_equal(image.view(np.uint32), single.view(np.uint32))
            print(f"world={world} {w}x{h} bounces={bounces} exchange={mode}: paths={paths} (expected {w * h}), "
                  f"image on rank 0 bit-identical to single GPU: {same}", flush=True)
            assert same and paths == w * h
        assert "p2p" in results or world == 1, "peer-memory exchange was not available"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Multi-rank host logic on CPU: world_size-2 gloo processes render disjoint tile sets (with the oracle as
the stand-in renderer), sum-reduce to rank 0, and must reproduce the single-rank image bit-exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _oracle as O
import rayfinder_b200 as rf
from rayfinder_b200 import distributed as rfd

W, H, SPP, BOUNCES = 80, 72, 2, 3


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_path: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pt = rf.PtFormat.loads(O.duck_pt_bytes())
        cam = rf.camera_to_array(rf.bvh_visualizer_camera(pt.bvh_nodes, W, H))
        orc = O.OracleRenderer(pt, W, H, cam, rf.sky_state(rf.Sky()), SPP, BOUNCES, rank=rank, world=world, threads=1)
        t = torch.from_numpy(orc.image)  # the rank's accumulation buffer (the oracle adds into it in place)
        exchange = rfd.HdrExchange(None, W, H, mode="auto", tensor=t)  # host tensor: the reduce, whatever mode is asked for
        assert exchange.mode == "nccl" and exchange.world == world and exchange.rank == rank
        owner = rfd.tile_owner(W, H, world)
        for _ in range(SPP):
            # progressive render: the exchange runs after EVERY frame and must not feed back into the accumulation
            orc.render()
            reduced = exchange()
            assert reduced is not t
            assert np.all(orc.image[owner != rank] == 0.0)  # non-owned pixels stay exactly zero on every rank, root included
        assert orc.stats()["paths"] == SPP * rfd.owned_pixel_count(W, H, rank, world)
        t = reduced
        exchange.close()
        with pytest.raises(ValueError):
            rfd.HdrExchange(None, W, H, mode="p2p", tensor=torch.from_numpy(orc.image))
        counters = torch.from_numpy(orc.counters.astype(np.int64))
        dist.all_reduce(counters)
        if rank == 0:
            np.savez(out_path, image=t.numpy(), counters=counters.numpy())
    finally:
        dist.destroy_process_group()


def test_tile_owner_partition_is_exact():
    for (w, h, world) in ((1920, 1080, 8), (3840, 2160, 8), (100, 70, 3), (31, 31, 2)):
        owner = rfd.tile_owner(w, h, world)
        counts = [rfd.owned_pixel_count(w, h, r, world) for r in range(world)]
        assert sum(counts) == w * h and owner.min() == 0 and owner.max() <= world - 1
    # load balance at the benchmark sizes: every rank within 5 % of the mean
    for (w, h) in ((1920, 1080), (3840, 2160)):
        counts = np.array([rfd.owned_pixel_count(w, h, r, 8) for r in range(8)])
        assert counts.max() / counts.mean() < 1.05


@pytest.mark.timeout(300)
def test_two_rank_reduce_is_bit_identical_to_single_rank(tmp_path):
    out = tmp_path / "reduced.npz"
    mp.spawn(_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    got = np.load(out)
    pt = rf.PtFormat.loads(O.duck_pt_bytes())
    cam = rf.camera_to_array(rf.bvh_visualizer_camera(pt.bvh_nodes, W, H))
    single = O.OracleRenderer(pt, W, H, cam, rf.sky_state(rf.Sky()), SPP, BOUNCES, threads=1)
    for _ in range(SPP):
        single.render()
    assert np.array_equal(got["image"].view(np.uint32), single.image.view(np.uint32))
    assert np.array_equal(got["counters"], single.counters.astype(np.int64))

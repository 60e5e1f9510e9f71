"""Multi-GPU plumbing of the render path: one process per GPU, frames partitioned by 32x32 pixel tile, and one
exchange step per frame that brings the HDR accumulation buffer together on rank 0 — either a sum-reduce (NCCL over
NVLink; gloo on CPU for the tests) or, fused into the accumulation kernel, direct stores of every owned pixel into rank
0's buffer over NVLink peer memory (`HdrExchange`, mode "p2p").

The reference is single-GPU; this is the only exchange step the path has.  Tiles are
disjoint and every non-owned pixel is exactly 0.0f, so the fp32 sum is exact and the reduced image is
bit-identical to a single-GPU frame.
"""
from __future__ import annotations

import numpy as np

TILE = 32


def tile_owner(width: int, height: int, world: int) -> np.ndarray:
    """(height, width) int array: owning rank of each pixel — tile (tx, ty) belongs to (tx + ty) % world,
    the rule of rf_renderer_set_tile_partition."""
    ty, tx = np.meshgrid(np.arange(height) // TILE, np.arange(width) // TILE, indexing="ij")
    return (tx + ty) % world


def owned_pixel_count(width: int, height: int, rank: int, world: int) -> int:
    return int((tile_owner(width, height, world) == rank).sum())


class _DeviceArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, num_floats: int):
        self.__cuda_array_interface__ = {"shape": (num_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def hdr_tensor(renderer, width: int, height: int):
    """The renderer's HDR sum buffer as a torch CUDA tensor (height, width, 4), no copy."""
    import torch

    t = torch.as_tensor(_DeviceArray(renderer.hdr_device_ptr(), width * height * 4), device="cuda")
    return t.view(height, width, 4)


def reduce_hdr(tensor, dst: int = 0, group=None):
    """Sum-reduce ``tensor`` to ``dst`` in place.  ``tensor`` must hold THIS FRAME's exchange copy, never a buffer that
    keeps accumulating across frames on ``dst`` (see HdrExchange: the root's own accumulation buffer would be summed with
    the other ranks' cumulative values again every frame)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if tensor.is_cuda and dist.get_backend(group) == "gloo":
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)  # gloo has no CUDA reduce; same sum
        else:
            dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


class HdrExchange:
    """The per-frame exchange step of the multi-GPU path, called right after ``renderer.render()`` on every rank; returns
    the full frame on the root (a tensor that stays valid until the next call).

    mode "nccl": one ``reduce(SUM)`` of the W x H x 16-byte HDR buffer to ``root`` (what BASELINE.json's north_star names).
                 OUT OF PLACE: every rank copies its accumulation buffer into an exchange tensor and that tensor is
                 reduced, so the root's accumulation buffer keeps holding only the root's own pixels — reducing in place
                 would add the other ranks' cumulative sums to it again on every frame of a progressive (spp > 1) render.
    mode "p2p":  the exchange is fused into the accumulation kernel — the root exports a double-buffered exchange target
                 through CUDA IPC, the other ranks map it, and every rank's ``k_accumulate`` stores each owned pixel's
                 accumulated value straight into half (frame & 1) of it, over NVLink peer memory; what is left per frame
                 is a 4-byte all-reduce that orders "all ranks have finished the frame" before the root uses the image
                 (and keeps the ranks in step).  16 bytes per owned pixel cross NVLink instead of W x H x 16 bytes of
                 mostly zeros going through a reduction.  The root may read frame N while the others already store
                 frame N + 1 (other half); include/rayfinder_b200.h has the ordering argument.
    mode "auto": "p2p" if every rank could map the root's buffer, else "nccl".
    Both produce the single-GPU image bit for bit on the root, also when called after every frame of a progressive
    render (tools/check_multigpu.py, tests/test_gpu_exchange.py)."""

    def __init__(self, renderer, width: int, height: int, mode: str = "auto", root: int = 0, group=None, tensor=None):
        """``tensor``: exchange this (height, width, 4) tensor instead of the renderer's device buffer — host tensors with the
        gloo backend in the CPU tests; only the reduce is possible then."""
        import torch
        import torch.distributed as dist

        self.renderer, self.root, self.group = renderer, root, group
        self.width, self.height = width, height
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.hdr = tensor if tensor is not None else hdr_tensor(renderer, width, height)  # this rank's accumulation buffer
        self.result = None  # nccl mode: the exchange tensor (allocated on first use)
        self._views = {}    # p2p mode, root: device pointer -> tensor view of that half of the exchange buffer
        self.mode = "nccl"
        if tensor is not None and mode == "p2p":
            raise ValueError("HdrExchange: the peer-memory exchange needs the renderer's device buffer")
        if self.world == 1 or mode == "nccl" or tensor is not None:
            return
        handle = [renderer.hdr_ipc_handle() if self.rank == root else None]
        dist.broadcast_object_list(handle, src=root, group=group)
        ok = torch.ones(1, dtype=torch.int32, device=self.hdr.device)
        if self.rank != root:
            try:
                renderer.set_hdr_peer(handle[0])
            except Exception:  # no peer access between the two devices: every rank falls back together
                ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            self.mode = "p2p"
            self._token = torch.zeros(1, dtype=torch.int32, device=self.hdr.device)
        else:
            renderer.set_hdr_peer(None)
            if mode == "p2p":
                raise RuntimeError("HdrExchange: peer-memory exchange requested but a rank could not map the root's HDR buffer")

    def _presented(self):
        import torch

        ptr = self.renderer.exchange_device_ptr()
        if ptr not in self._views:
            t = torch.as_tensor(_DeviceArray(ptr, self.width * self.height * 4), device=self.hdr.device)
            self._views[ptr] = t.view(self.height, self.width, 4)
        return self._views[ptr]

    def __call__(self):
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return self.hdr
        if self.mode == "p2p":
            dist.all_reduce(self._token, group=self.group)  # stream-ordered after this rank's kernels: the frame barrier
            return self._presented() if self.rank == self.root else self.hdr
        if self.result is None:
            self.result = torch.empty_like(self.hdr)
        self.result.copy_(self.hdr)
        reduce_hdr(self.result, dst=self.root, group=self.group)
        return self.result

    def close(self):
        if self.mode == "p2p":
            import torch.distributed as dist

            self._views.clear()
            self.renderer.set_hdr_peer(None)  # peers unmap the root's buffer, the root presents its local image again
            dist.barrier(group=self.group)    # ... before the root may free it
            self.mode = "nccl"

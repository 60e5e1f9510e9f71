"""Multi-GPU plumbing of the render path: one process per GPU, frames partitioned by 32x32 pixel tile,
one sum-reduce of the HDR accumulation buffer to rank 0 (NCCL over NVLink; gloo on CPU for the tests).

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
    """The per-frame exchange step: sum-reduce the HDR buffer to `dst` (in place on `dst`)."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor

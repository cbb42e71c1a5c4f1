"""One process per GPU: frame sharding and the final gather of radiance rows to rank 0.

The reference has no multi-device path; pixels are independent, so the frame is split
into interleaved scanline tiles (the reference's own 4-scanline batch, ocltracer.go:214-223, is the
tile) and the only exchange is one gather of each rank's rows to rank 0 -- NCCL over NVLink on
GPUs, gloo on CPU for the host-logic tests.  `torch.distributed` is plumbing only.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import trace as T


class _DeviceBuffer:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def framebuffer_tensor(ctx: "T.Context", local_index: int = 0) -> torch.Tensor:
    """The packed rows a context rendered on one of its devices, as a CUDA float64 tensor view."""
    ptr, n, dev = ctx.device_framebuffer(local_index)
    if n == 0:
        return torch.empty(0, dtype=torch.float64, device=f"cuda:{dev}")
    return torch.as_tensor(_DeviceBuffer(ptr, n), device=f"cuda:{dev}")


def all_shard_rows(height: int, world: int, rows_per_tile: int = 0) -> List[np.ndarray]:
    return [T.plan_rows(height, r, world, rows_per_tile) for r in range(world)]


def gather_frame(local: torch.Tensor, height: int, width: int, rows_per_tile: int = 0, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather every rank's packed rows ([n_rows_r * width * 4] float64) into the full frame on `dst`.

    Returns the [height, width, 4] frame on rank `dst` (same device as `local`), None elsewhere."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    rows = all_shard_rows(height, world, rows_per_tile)
    row_len = width * 4
    if local.numel() != len(rows[rank]) * row_len:
        raise ValueError(f"rank {rank}: expected {len(rows[rank])} rows, got {local.numel() // row_len}")
    if world == 1:
        return local.view(height, width, 4)
    max_rows = max(len(r) for r in rows)
    padded = local
    if len(rows[rank]) < max_rows:
        padded = torch.zeros(max_rows * row_len, dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    frame = torch.empty(height, row_len, dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = torch.as_tensor(rows[r], dtype=torch.long, device=local.device)
        frame.index_copy_(0, idx, bufs[r].view(max_rows, row_len)[: len(rows[r])])
    return frame.view(height, width, 4)

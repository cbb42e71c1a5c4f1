"""One process per GPU: frame sharding and the final gather of radiance rows to rank 0.

The reference has no multi-device path; pixels are independent, so the frame is split
into interleaved scanline tiles (the reference's own 4-scanline batch, ocltracer.go:214-223, is the
tile) and the only exchange is the gather of each rank's rows on rank 0.

On GPUs that gather is FUSED into the trace kernel: rank 0 owns a whole-frame buffer (`trace.Frame`), the
other ranks map it through CUDA IPC (`FrameExchange`), and every rank's kernel stores its finished pixels
straight into it -- NVLink peer stores from the kernel epilogue, no copy kernel, no staging, no collective
on the data path; `torch.distributed` only carries the 64-byte handle and the "all shards traced" barrier.
`gather_frame` is the plain collective (NCCL or gloo) kept for hosts without peer access and for the CPU
tests of the host logic.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import trace as T


class _DeviceBuffer:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def framebuffer_tensor(ctx: "T.Context", local_index: int = 0) -> torch.Tensor:
    """The packed rows a context rendered on one of its devices, as a CUDA float64 tensor view."""
    ptr, n, dev = ctx.device_framebuffer(local_index)
    if n == 0:
        return torch.empty(0, dtype=torch.float64, device=f"cuda:{dev}")
    return torch.as_tensor(_DeviceBuffer(ptr, n), device=f"cuda:{dev}")


class FrameExchange:
    """A frame on rank `dst`'s GPU that every rank renders into directly (see the module docstring).

    Create it once per (size, format) and attach it to any number of successive contexts with
    ``ctx.set_frame(ex.frame)``; after ``ctx.trace()`` on every rank and ``ex.barrier()``, rank `dst` holds the
    complete image: ``ex.tensor()`` (device view) or ``ex.frame.read()`` (host copy)."""

    def __init__(self, width: int, height: int, device: int, dst: int = 0, fmt: int = T.FRAME_F64):
        self.width, self.height, self.dst, self.format = width, height, dst, fmt
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        box: List[Optional[bytes]] = [None]
        if self.rank == dst:
            self.frame = T.Frame(device, width, height, fmt)
            box[0] = self.frame.export() if self.world > 1 else None
        if self.world > 1:
            dist.broadcast_object_list(box, src=dst)
            if self.rank != dst:
                self.frame = T.Frame.attach(device, box[0], width, height, fmt)

    def barrier(self) -> None:
        """Orders "every rank's trace kernel has finished" (ctx.trace() waits for its own) before the read."""
        if self.world > 1:
            dist.barrier()

    def tensor(self) -> Optional[torch.Tensor]:
        """[H, W, 4] device view of the frame on rank `dst`, None elsewhere."""
        if self.rank != self.dst:
            return None
        ptr, nbytes, dev = self.frame.device_pointer()
        if self.format == T.FRAME_F32:
            t = torch.as_tensor(_DeviceBuffer(ptr, nbytes // 4, "<f4"), device=f"cuda:{dev}")
        else:
            t = torch.as_tensor(_DeviceBuffer(ptr, nbytes // 8), device=f"cuda:{dev}")
        return t.view(self.height, self.width, 4)

    def close(self) -> None:
        self.frame.close()


def all_shard_rows(height: int, world: int, rows_per_tile: int = 0) -> List[np.ndarray]:
    return [T.plan_rows(height, r, world, rows_per_tile) for r in range(world)]


_ROW_INDEX_CACHE: Dict[Tuple, List[torch.Tensor]] = {}


def _row_indices(height: int, world: int, rows_per_tile: int, device) -> List[torch.Tensor]:
    """Row-index tensors of every shard on `device`, built once per (frame, world, device)."""
    key = (height, world, rows_per_tile, str(device))
    if key not in _ROW_INDEX_CACHE:
        _ROW_INDEX_CACHE[key] = [torch.as_tensor(r, dtype=torch.long, device=device) for r in all_shard_rows(height, world, rows_per_tile)]
    return _ROW_INDEX_CACHE[key]


def gather_frame(local: torch.Tensor, height: int, width: int, rows_per_tile: int = 0, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather every rank's packed rows ([n_rows_r * width * 4] float64) into the full frame on `dst` with the
    backend's gather collective.  Returns the [height, width, 4] frame on rank `dst`, None elsewhere."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    idx = _row_indices(height, world, rows_per_tile, local.device)
    row_len = width * 4
    if local.numel() != len(idx[rank]) * row_len:
        raise ValueError(f"rank {rank}: expected {len(idx[rank])} rows, got {local.numel() // row_len}")
    if world == 1:
        return local.view(height, width, 4)
    max_rows = max(len(r) for r in idx)
    padded = local
    if len(idx[rank]) < max_rows:
        padded = torch.zeros(max_rows * row_len, dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    frame = torch.empty(height, row_len, dtype=local.dtype, device=local.device)
    for r in range(world):
        frame.index_copy_(0, idx[r], bufs[r].view(max_rows, row_len)[: len(idx[r])])
    return frame.view(height, width, 4)

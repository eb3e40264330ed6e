"""Tile scheduler over the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

The reference scatters row bands over slaves with HTTP and gathers `ImageSlice` callbacks
(ray-tracer-controller/src/main.rs:47-75,79-93,109-115).  Here the frame is cut into 8x4-pixel tiles,
tile t belongs to rank `tile_owner(t, ranks)` (a rotating interleave, so sky-heavy and object-heavy
regions spread evenly), the scene is replicated on every GPU like the reference replicates it in every
request body, and the frame is assembled on rank 0 in one of two ways:

  * "p2p"  (fused render + gather): rank 0 owns the frame in its HBM and exports a CUDA IPC handle; every
    other rank maps it and its render kernel stores finished tiles straight into rank 0's memory over
    NVLink.  One barrier ends the frame; there is no separate collective and no staging copy.
  * "nccl" (baseline): every rank renders its tiles into a zeroed local frame; an NCCL reduce(MAX) over the
    uint8 frames to rank 0 assembles them (tiles are disjoint, everything else is zero).

Pixels draw from per-pixel streams, so the image does not depend on the partition.
"""
from __future__ import annotations

import numpy as np

TILE_W, TILE_H = 8, 4  # must match csrc/rt_device.cuh


def tile_grid(width: int, height: int):
    return (width + TILE_W - 1) // TILE_W, (height + TILE_H - 1) // TILE_H


def tile_owner(tile: int, ranks: int) -> int:
    """Rank that renders global tile index `tile` (row-major over the tile grid).

    Inverse of the kernel's ticket map g = k*ranks + (rank + k) % ranks  (csrc/rt_kernels.cu)."""
    k, e = divmod(tile, ranks)
    return (e - k) % ranks


def owner_map(width: int, height: int, ranks: int) -> np.ndarray:
    """int32 (height, width): owning rank of every pixel."""
    tx, ty = tile_grid(width, height)
    t = np.arange(tx * ty, dtype=np.int64)
    own = ((t % ranks) - (t // ranks)) % ranks
    own = own.reshape(ty, tx).astype(np.int32)
    return np.repeat(np.repeat(own, TILE_H, axis=0), TILE_W, axis=1)[:height, :width]


def tiles_of_rank(width: int, height: int, rank: int, ranks: int) -> np.ndarray:
    tx, ty = tile_grid(width, height)
    t = np.arange(tx * ty, dtype=np.int64)
    return t[((t % ranks) - (t // ranks)) % ranks == rank]


def assemble_reduce(frame, dst: int = 0, group=None):
    """Assemble disjoint, zero-padded per-rank frames on `dst` with reduce(MAX).  Works on any backend
    (NCCL on GPU tensors; gloo on CPU tensors in the tests)."""
    import torch.distributed as dist

    dist.reduce(frame, dst=dst, op=dist.ReduceOp.MAX, group=group)
    return frame


class FrameScheduler:
    """Per-rank driver of a multi-GPU frame.  `ctx` is this rank's rt_b200.Context."""

    def __init__(self, ctx, rank: int, world_size: int, mode: str = "p2p"):
        import torch

        self.ctx, self.rank, self.world = ctx, rank, world_size
        self.mode = mode
        self.torch = torch
        self._frame_ptr = None
        self._frame_bytes = 0
        self._local = None
        self._owns = False

    # -- frame buffer ------------------------------------------------------------------------------
    def setup(self, width: int, height: int):
        import torch.distributed as dist

        torch = self.torch
        nbytes = width * height * 3
        self.close()
        self._frame_bytes = nbytes
        if self.world == 1:
            self._frame_ptr, _ = self.ctx.frame_alloc(nbytes)
            self._owns = True
            return self
        if self.mode == "p2p":
            obj = [None]
            if self.rank == 0:
                self._frame_ptr, handle = self.ctx.frame_alloc(nbytes)
                self._owns = True
                obj = [handle]
            dist.broadcast_object_list(obj, src=0)
            if self.rank != 0:
                self._frame_ptr = self.ctx.frame_open(obj[0])
        elif self.mode == "nccl":
            self._local = torch.zeros(nbytes, dtype=torch.uint8, device=f"cuda:{self.ctx.device}")
            self._frame_ptr = self._local.data_ptr()
        else:
            raise ValueError(f"unknown mode {self.mode!r}")
        return self

    def close(self):
        if self._frame_ptr is not None and self._local is None:
            if self._owns:
                self.ctx.frame_free(self._frame_ptr)
            else:
                self.ctx.frame_close(self._frame_ptr)
        self._frame_ptr, self._local, self._owns = None, None, False

    # -- one frame -----------------------------------------------------------------------------------
    def render(self, scene, params, want_stats: bool = False):
        """Render this rank's tiles and complete the frame on rank 0 (device resident).  Collective."""
        import torch.distributed as dist

        torch = self.torch
        if self.world > 1 and self.mode == "nccl":
            self._local.zero_()
            torch.cuda.current_stream().synchronize()
        st = self.ctx.render_tiles_device(scene, params, self.rank, self.world, self._frame_ptr, sync=True,
                                          want_stats=True)
        if self.world > 1:
            if self.mode == "nccl":
                assemble_reduce(self._local, 0)
                torch.cuda.current_stream().synchronize()
            else:
                dist.barrier()
        return st if want_stats else None

    def download(self, out: np.ndarray) -> np.ndarray:
        """Rank 0: copy the assembled frame to host memory."""
        assert self.rank == 0
        return self.ctx.frame_download(self._frame_ptr, out)

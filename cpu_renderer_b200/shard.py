"""Multi-GPU decomposition of the hot path: one process per GPU.

The path shards without any data-path collective: rows of the reference's row loop are
independent (projekt.cpp:198), so a GPU can own a band of screen rows (config C4), and frames are
independent (config C5).  The only exchange is the gather of finished band / frame images
(torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def band_rows(height: int, world: int, rank: int, tile_h: int = 32):
    """Rows [first, first+rows) owned by ``rank``: bands are whole tile rows so that no screen tile
    straddles two GPUs; the remainder goes to the last ranks one tile row at a time."""
    tile_rows = (height + tile_h - 1) // tile_h
    base, extra = divmod(tile_rows, world)
    # the first (world - extra) ranks get `base` tile rows, the rest base + 1
    starts = []
    at = 0
    for r in range(world):
        n = base + (1 if r >= world - extra else 0)
        starts.append((at, n))
        at += n
    t0, n = starts[rank]
    first = min(t0 * tile_h, height)
    last = min((t0 + n) * tile_h, height)
    return first, last - first


def frame_range(nframes: int, world: int, rank: int) -> range:
    """Contiguous block of frames (views) rendered by ``rank`` (C5: 256 views over 8 GPUs)."""
    base, extra = divmod(nframes, world)
    first = rank * base + min(rank, extra)
    return range(first, first + base + (1 if rank < extra else 0))


def gather_bands(local: torch.Tensor, height: int, world: int, rank: int, tile_h: int = 32, dst: int = 0):
    """Gather row bands ([rows_r, W] tensors, rows as given by band_rows) into the full [H, W]
    image on ``dst``.  Bands are padded to a common height so one dist.gather moves everything."""
    if world == 1:
        return local
    sizes = [band_rows(height, world, r, tile_h) for r in range(world)]
    max_rows = max(n for _, n in sizes)
    padded = local
    if local.shape[0] < max_rows:
        padded = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded.contiguous(), parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat([parts[r][:sizes[r][1]] for r in range(world)], dim=0)


def gather_frames(local: torch.Tensor, world: int, rank: int, dst: int = 0):
    """Gather equally many frames per rank ([frames_r, H, W]) into [frames, H, W] on ``dst``."""
    if world == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local.contiguous(), parts, dst=dst)
    return torch.cat(parts, dim=0) if rank == dst else None

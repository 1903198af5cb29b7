"""A 64-bit image hash that can be computed where the image lives (GPU tensor or numpy array).

Word-wise FNV-1a over a 16384 x 16384 image is a 268 M-step sequential chain; this hash keeps FNV-1a-64
for the fold but feeds it one 64-bit digest per ROW, and a row digest is a position-weighted sum that
any parallel reduction can produce:

    row(y)  = sum_x (pixel[y, x] + 1) * (x * 0x9E3779B97F4A7C15 + 1)      (mod 2^64)
    image   = FNV-1a-64 over the words  lo32(row(0)), hi32(row(0)), lo32(row(1)), ...

Used by bench.py (`image_fnv`): the frame a rank rendered -- for the row-band split the image rank 0
gathered over NCCL -- against the same hash of the CPU oracle's frame (tools/make_image_hashes.py).
"""
from __future__ import annotations

import numpy as np

_K = 0x9E3779B97F4A7C15
_FNV_OFFSET, _FNV_PRIME = 0xcbf29ce484222325, 0x100000001b3


def _fold(rows_u64: np.ndarray) -> str:
    h = _FNV_OFFSET
    for r in rows_u64.tolist():
        for w in (r & 0xFFFFFFFF, r >> 32):
            h = ((h ^ w) * _FNV_PRIME) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def image_fnv_numpy(image: np.ndarray) -> str:
    """image: [H, W] uint32 (or float32, hashed by its bit pattern)."""
    a = np.ascontiguousarray(image).view(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        wts = np.arange(a.shape[1], dtype=np.uint64) * np.uint64(_K) + np.uint64(1)
        rows = ((a + np.uint64(1)) * wts[None, :]).sum(axis=1, dtype=np.uint64)
    return _fold(rows)


def image_fnv_torch(image, width: int) -> str:
    """image: [H, >= width] int32 / float32 CUDA (or CPU) tensor; columns beyond `width` are padding."""
    import torch
    a = image[:, :width]
    if a.dtype != torch.int32:
        a = a.contiguous().view(torch.int32)
    signed_k = _K - (1 << 64)                      # the same multiplier as a wrapped int64
    wts = torch.arange(width, dtype=torch.int64, device=a.device) * signed_k + 1
    rows = torch.empty(a.shape[0], dtype=torch.int64, device=a.device)
    step = max(1, (1 << 26) // max(width, 1))      # bound the int64 temporaries (a 16K^2 image is 2 GiB as int64)
    for y in range(0, a.shape[0], step):
        blk = (a[y:y + step].to(torch.int64) & 0xFFFFFFFF) + 1
        rows[y:y + step] = (blk * wts[None, :]).sum(dim=1)
    return _fold(rows.cpu().numpy().view(np.uint64))

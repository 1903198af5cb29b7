"""Multi-GPU decomposition of the hot path: one process per GPU.

The path shards without any data-path collective: rows of the reference's row loop are
independent (projekt.cpp:198), so a GPU can own a band of screen rows (config C4), and frames are
independent (config C5).  The only exchange is the gather of finished band / frame images: either a collective after the frame
(torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests), or fused into the raster
kernel's tile write-back (FusedGather below: the kernel stores every finished tile a second time,
straight into the assembling GPU's memory over NVLink, and only a barrier follows the frame).
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


class _DevicePointer:
    """Lets torch.as_tensor view raw device memory (a b200r_peer_alloc allocation) without copying."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class FusedGather:
    """The gather of the finished images fused into the raster kernel (b200r_set_gather_target).

    ``dst`` allocates ONE image of ``slots`` whole screens ([slots, height, wpad] u32 colour, optionally
    f32 depth) with b200r_peer_alloc and sends its CUDA IPC handles to the other ranks, which map it
    (b200r_peer_open).  Every rank then points its renderer's gather target at the screen it
    contributes to -- the same slot for row bands of one frame, slot = rank for frame-parallel work --
    and from then on each b200r_render_device call also stores its tiles there.  ``finish()`` is the
    stream-ordered barrier after which ``dst`` may read the assembled image(s)."""

    def __init__(self, renderer, api, height: int, width: int, wpad: int, slots: int, world: int, rank: int,
                 device, dst: int = 0, with_depth: bool = False):
        self.r, self.api, self.rank, self.dst, self.world = renderer, api, rank, dst, world
        self.h, self.w, self.wpad, self.slots = height, width, wpad, slots
        nbytes = slots * height * wpad * 4
        handles = [None, None]
        self._owned, self._opened = [], []
        if rank == dst:
            self.color_ptr, hc = renderer.peer_alloc(nbytes)
            self._owned.append(self.color_ptr)
            handles[0] = hc
            if with_depth:
                self.depth_ptr, hd = renderer.peer_alloc(nbytes)
                self._owned.append(self.depth_ptr)
                handles[1] = hd
        if world > 1:
            dist.broadcast_object_list(handles, src=dst)
        if rank != dst:
            self.color_ptr = renderer.peer_open(handles[0])
            self._opened.append(self.color_ptr)
            if handles[1] is not None:
                self.depth_ptr = renderer.peer_open(handles[1])
                self._opened.append(self.depth_ptr)
        if not with_depth:
            self.depth_ptr = 0
        self.color = self.depth = None
        if rank == dst:
            self.color = torch.as_tensor(_DevicePointer(self.color_ptr, (slots, height, wpad), "<i4"), device=device)
            if with_depth:
                self.depth = torch.as_tensor(_DevicePointer(self.depth_ptr, (slots, height, wpad), "<f4"), device=device)
        self._token = torch.zeros(1, dtype=torch.int32, device=device)

    def select(self, slot: int):
        """Mirror this rank's frames into screen ``slot`` of the assembled image."""
        off = slot * self.h * self.wpad * 4
        t = self.api.device_target(self.color_ptr + off, (self.depth_ptr + off) if self.depth_ptr else None,
                                   self.w, self.h, self.wpad * 4, self.wpad, 0, self.h)
        self.r.set_gather_target(t)

    def finish(self):
        """Stream-ordered barrier: enqueued behind this rank's frame on the current stream; when it has
        completed on ``dst``, every rank's kernels -- and with them their stores into the image -- have."""
        if self.world > 1:
            dist.all_reduce(self._token)

    def close(self):
        self.r.set_gather_target(None)
        self.color = self.depth = None
        for p in self._opened + self._owned:
            self.r.peer_release(p)
        self._opened, self._owned = [], []

"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo process groups.
Each rank produces its rows / frames with the CPU oracle; the gather must reassemble the
single-process image exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from cpu_renderer_b200 import scene as sc
from cpu_renderer_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_band_rows_partition_every_row_once():
    for height in (1080, 2160, 16384, 200, 33):
        for world in (1, 2, 3, 4, 8):
            for tile_h in (16, 32):
                at = 0
                for r in range(world):
                    first, rows = shard.band_rows(height, world, r, tile_h)
                    assert first == at and rows >= 0
                    assert first % tile_h == 0 or first == height
                    at += rows
                assert at == height


def test_frame_range_partition():
    for n in (256, 7, 1):
        for world in (1, 2, 8):
            got = [f for r in range(world) for f in shard.frame_range(n, world, r)]
            assert got == list(range(n))


def _band_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = sc.triangle_soup("bands", 0x44, 4000, 640, 360, 2.0, 40.0)
    full = ol.oracle_render(s)            # rows are independent: a band is a slice of the frame
    first, rows = shard.band_rows(s.height, world, rank, 32)
    mine = torch.from_numpy(full["color"][first:first + rows].astype(np.int64))
    img = shard.gather_bands(mine, s.height, world, rank, 32, dst=0)
    if rank == 0:
        out.put(bool(np.array_equal(img.numpy().astype(np.uint32), full["color"])))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_gather_reassembles_the_frame(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def _frame_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = []
    for f in shard.frame_range(4, world, rank):       # every frame = its own seed, as in bench.py
        s = sc.triangle_soup("f", 0x900 + f, 500, 160, 120, 2.0, 20.0)
        frames.append(ol.oracle_render(s)["color"].astype(np.int64))
    got = shard.gather_frames(torch.from_numpy(np.stack(frames)), world, rank, dst=0)
    if rank == 0:
        want = np.stack([ol.oracle_render(sc.triangle_soup("f", 0x900 + f, 500, 160, 120, 2.0, 20.0))["color"]
                         for f in range(4)])
        out.put(bool(np.array_equal(got.numpy().astype(np.uint32), want)))
    dist.destroy_process_group()


def test_frame_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_frame_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True

"""The fused gather (b200r_set_gather_target): bands rendered on separate contexts /
processes mirror their tiles into ONE image of the whole screen, which must equal the oracle's
frame bit for bit -- including tiles nothing was drawn into and pre-existing target contents.

The multi-GPU bench uses this over NVLink; here the same code path runs on one GPU: a mirror in
plain device memory, and a mirror allocated with b200r_peer_alloc that a SECOND PROCESS maps with
b200r_peer_open (CUDA IPC) and renders its band into.
"""
import os
import subprocess
import sys
from dataclasses import replace

import numpy as np
import pytest

import oracle_lib as ol
from cpu_renderer_b200 import api
from cpu_renderer_b200 import scene as sc
from cpu_renderer_b200 import shard
from cpu_renderer_b200.api import Renderer

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _upload(s, dev):
    import torch
    return [torch.from_numpy(a).to(dev) for a in (s.positions, s.colors, s.normals)]


def _render_band(r, s, bufs, first, rows, wpad, dev, tile=(0, 0), pre=None):
    """One band through b200r_render_device; returns the band's own colour / depth tensors."""
    import torch
    color = torch.full((rows, wpad), s.clear_color, dtype=torch.int32, device=dev)
    depth = torch.full((rows, wpad), s.clear_depth, dtype=torch.float32, device=dev)
    if pre is not None:
        color[:, :s.width] = torch.from_numpy(pre[0][first:first + rows].view(np.int32)).to(dev)
        depth[:, :s.width] = torch.from_numpy(pre[1][first:first + rows]).to(dev)
    mesh = api.device_mesh(bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), s.triangle_count,
                           api.v3(*s.object_p), 0, None, None)
    cmd, keep = api.make_commands(s)
    tgt = api.device_target(color.data_ptr(), depth.data_ptr(), s.width, s.height, wpad * 4, wpad, first, rows)
    r.set_tile(*tile)
    r.render_device([mesh], cmd, tgt)
    r.sync()
    return color, depth


@pytest.mark.parametrize("width,wpad_to,tile", [(1280, 64, (0, 0)), (1280, 64, (64, 32)), (1001, 1, (128, 8)), (700, 64, (64, 16))])
def test_bands_mirrored_into_one_image_equal_the_oracle_frame(width, wpad_to, tile):
    import torch
    dev = torch.device("cuda", 0)
    # sparse scene: most tiles of the lower bands stay empty and must be copied through
    s = sc.triangle_soup("gather", 0x71, 6000, width, 360, 2.0, 30.0)
    keep = (s.positions[:, 1].reshape(-1, 3) < 0.0).all(axis=1).repeat(3)      # one half of the screen stays empty
    s = replace(s, positions=np.ascontiguousarray(s.positions[keep]), colors=np.ascontiguousarray(s.colors[keep]),
                normals=np.ascontiguousarray(s.normals[keep]), uvs=np.ascontiguousarray(s.uvs[keep]))
    rng = np.random.default_rng(5)
    pre_c = rng.integers(0, 2**32, size=(s.height, s.width), dtype=np.uint64).astype(np.uint32)
    pre_z = np.where(rng.random((s.height, s.width)) < 0.5, np.float32(s.clear_depth), np.float32(1.0e9)).astype(np.float32)
    want = ol.oracle_render(s, targets=(pre_c.copy(), pre_z.copy(), None))
    wpad = (width + wpad_to - 1) // wpad_to * wpad_to
    g_color = torch.full((s.height, wpad), 0x55555555, dtype=torch.int32, device=dev)
    g_depth = torch.full((s.height, wpad), -7.0, dtype=torch.float32, device=dev)
    gather = api.device_target(g_color.data_ptr(), g_depth.data_ptr(), s.width, s.height, wpad * 4, wpad, 0, s.height)
    world = 3
    ctxs = [Renderer(0) for _ in range(world)]          # one context per band, as one per GPU
    try:
        for rank, r in enumerate(ctxs):
            r.set_gather_target(gather)
            bufs = _upload(s, dev)
            first, rows = shard.band_rows(s.height, world, rank, 32)
            c, z = _render_band(r, s, bufs, first, rows, wpad, dev, tile, pre=(pre_c, pre_z))
            # the band itself is unaffected by the mirror
            assert np.array_equal(c[:, :width].cpu().numpy().view(np.uint32), want["color"][first:first + rows])
            assert np.array_equal(z[:, :width].cpu().numpy().view(np.uint32), want["z"][first:first + rows].view(np.uint32))
    finally:
        for r in ctxs:
            r.close()
    got_c = g_color[:, :width].cpu().numpy().view(np.uint32)
    got_z = g_depth[:, :width].cpu().numpy()
    assert np.array_equal(got_c, want["color"])
    assert np.array_equal(got_z.view(np.uint32), want["z"].view(np.uint32))
    if wpad > width:                                     # row padding of the gathered image is never written
        assert (g_color[:, width:] == 0x55555555).all() and (g_depth[:, width:] == -7.0).all()


def test_gather_target_validation_and_switch_off():
    import torch
    dev = torch.device("cuda", 0)
    s = sc.triangle_soup("gather", 0x72, 500, 640, 360, 2.0, 30.0)
    r = Renderer(0)
    try:
        g = torch.zeros((s.height, 640), dtype=torch.int32, device=dev)
        bad = api.device_target(g.data_ptr(), None, 320, s.height, 640 * 4, 640, 0, s.height)     # another screen
        r.set_gather_target(bad)
        bufs = _upload(s, dev)
        with pytest.raises(api.B200RasterError):
            _render_band(r, s, bufs, 0, s.height, 640, dev)
        with pytest.raises(api.B200RasterError):
            r.set_gather_target(api.device_target(None, None, 640, 360, 640 * 4, 640, 0, 360))
        r.set_gather_target(None)                        # off: renders as before, the image is not written
        _render_band(r, s, bufs, 0, s.height, 640, dev)
        assert int(g.abs().sum().item()) == 0
        # colour only (Depth == 0) is allowed
        r.set_gather_target(api.device_target(g.data_ptr(), None, 640, s.height, 640 * 4, 640, 0, s.height))
        c, _ = _render_band(r, s, bufs, 0, s.height, 640, dev)
        assert torch.equal(c, g)
    finally:
        r.close()


CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from cpu_renderer_b200 import api, scene as sc, shard
from cpu_renderer_b200.api import Renderer
hc, hd = bytes.fromhex(sys.argv[1]), bytes.fromhex(sys.argv[2])
dev = torch.device("cuda", 0)
s = sc.triangle_soup("ipc", 0x73, 20000, 1280, 720, 2.0, 40.0)
r = Renderer(0)
pc, pd = r.peer_open(hc), r.peer_open(hd)
wpad = 1280
r.set_gather_target(api.device_target(pc, pd, s.width, s.height, wpad * 4, wpad, 0, s.height))
first, rows = shard.band_rows(s.height, 2, 1, 32)
bufs = [torch.from_numpy(a).to(dev) for a in (s.positions, s.colors, s.normals)]
color = torch.full((rows, wpad), s.clear_color, dtype=torch.int32, device=dev)
depth = torch.full((rows, wpad), s.clear_depth, dtype=torch.float32, device=dev)
mesh = api.device_mesh(bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), s.triangle_count, api.v3(*s.object_p), 0, None, None)
cmd, keep = api.make_commands(s)
r.render_device([mesh], cmd, api.device_target(color.data_ptr(), depth.data_ptr(), s.width, s.height, wpad * 4, wpad, first, rows))
r.sync()
r.peer_release(pc); r.peer_release(pd)
r.close()
print("child done")
"""


def test_second_process_renders_its_band_into_peer_memory():
    """CUDA IPC end to end on one GPU: this process allocates the image and renders band 0, a child
    process maps it and renders band 1 -- what ranks 0 and 1 of the multi-GPU bench do."""
    import torch
    dev = torch.device("cuda", 0)
    s = sc.triangle_soup("ipc", 0x73, 20000, 1280, 720, 2.0, 40.0)
    want = ol.oracle_render(s)
    wpad = 1280
    r = Renderer(0)
    try:
        nbytes = s.height * wpad * 4
        pc, hc = r.peer_alloc(nbytes)
        pd, hd = r.peer_alloc(nbytes)
        g_color = torch.as_tensor(shard._DevicePointer(pc, (s.height, wpad), "<i4"), device=dev)
        g_depth = torch.as_tensor(shard._DevicePointer(pd, (s.height, wpad), "<f4"), device=dev)
        g_color.fill_(0); g_depth.fill_(0.0)
        torch.cuda.synchronize()
        out = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT), hc.hex(), hd.hex()],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "child done" in out.stdout, out.stderr[-2000:]
        r.set_gather_target(api.device_target(pc, pd, s.width, s.height, wpad * 4, wpad, 0, s.height))
        first, rows = shard.band_rows(s.height, 2, 0, 32)
        _render_band(r, s, _upload(s, dev), first, rows, wpad, dev)
        got_c = g_color.cpu().numpy().view(np.uint32)
        got_z = g_depth.cpu().numpy()
        assert np.array_equal(got_c, want["color"])
        assert np.array_equal(got_z.view(np.uint32), want["z"].view(np.uint32))
        del g_color, g_depth
        r.peer_release(pc); r.peer_release(pd)
    finally:
        r.close()

"""Two contexts on two devices in ONE process (include/b200_raster.h: "one context per GPU").

The raster kernel's launch configuration (cudaFuncAttributeMaxDynamicSharedMemorySize, occupancy) is per
device; the 128x32 tile needs 80 KB of dynamic shared memory, above the 48 KB default, so a context on a
second device only works if that device was configured too.  Also checks that no entry point leaves the
calling thread on another device than it found it on.  Skipped on a single-GPU box."""
import numpy as np
import pytest

import oracle_lib as ol
from cpu_renderer_b200 import scene as sc
from cpu_renderer_b200.api import Renderer

pytestmark = pytest.mark.gpu


def test_two_contexts_on_two_devices_and_the_callers_device_is_restored():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = sc.triangle_soup("two", 0x2D, 20_000, 1280, 720, 2.0, 60.0)
    want = ol.oracle_render(s)
    torch.cuda.set_device(0)
    r0, r1 = Renderer(0), Renderer(1)
    assert torch.cuda.current_device() == 0
    try:
        for tile in ((128, 32), (64, 32), (256, 8)):
            for r in (r1, r0, r1):
                r.set_tile(*tile)
                color, z, _ = ol.new_targets(s)
                r.render_scene_host(s, color, z)
                assert torch.cuda.current_device() == 0          # the caller's device, not the context's
                assert np.array_equal(z.view(np.uint32), want["z"].view(np.uint32)), tile
                assert np.array_equal(color, want["color"]), tile
    finally:
        r0.close(); r1.close()
    assert torch.cuda.current_device() == 0

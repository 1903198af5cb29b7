"""GPU parity proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Coverage mask and depth bit-exact; colour bit-exact too (the stated bar is
+-1 LSB per 8-bit channel, the arithmetic is identical so 0 is expected and asserted)."""
import numpy as np
import pytest

import oracle_lib as ol
from cpu_renderer_b200 import scene as sc
from cpu_renderer_b200.api import Renderer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer():
    r = Renderer(0)
    yield r
    r.close()


def compare(renderer, scene, splits=None, tile=None):
    o = ol.oracle_render(scene)
    color, z, _ = ol.new_targets(scene)
    if tile:
        renderer.set_tile(*tile)
    renderer.render_scene_host(scene, color, z, splits=splits)
    zdiff = int((o["z"].view(np.uint32) != z.view(np.uint32)).sum())
    cov_o = o["z"] != np.float32(scene.clear_depth)
    cov_g = z != np.float32(scene.clear_depth)
    covdiff = int((cov_o != cov_g).sum())
    cdiff = int((o["color"] != color).sum())
    ch = np.abs(o["color"].view(np.uint8).astype(np.int16) - color.view(np.uint8).astype(np.int16))
    return dict(zdiff=zdiff, covdiff=covdiff, cdiff=cdiff, maxlsb=int(ch.max()), stats=o["stats"])


@pytest.mark.parametrize("tile", [(64, 32), (32, 32), (128, 16), (64, 16), (128, 32)])
def test_small_triangles_1080p(renderer, tile):
    s = sc.triangle_soup("small", 0xB2000002, 100_000, 1920, 1080, 1.5, 4.0)
    r = compare(renderer, s, tile=tile)
    assert r["covdiff"] == 0 and r["zdiff"] == 0 and r["cdiff"] == 0, r


@pytest.mark.parametrize("tile", [(64, 32), (128, 16)])
def test_large_overlapping(renderer, tile):
    s = sc.triangle_soup("large", 0xB2000003, 4000, 1920, 1080, 32.0, 96.0)
    r = compare(renderer, s, tile=tile)
    assert r["covdiff"] == 0 and r["zdiff"] == 0 and r["cdiff"] == 0, r

"""The oracle port against the verbatim reference build (oracle/_ref/libprojekt_ref.so), live.
Skipped where neither the prebuilt library nor /root/reference exists."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from cpu_renderer_b200 import api
from cpu_renderer_b200 import scene as sc

pytestmark = pytest.mark.skipif(not ol.ref_available(), reason="no verbatim reference build")


def test_struct_sizes_match_the_reference_build():
    lib = ol.ref()
    # projekt.h:2-98 + the Appendix-A structs; SURVEY.md Appendix C, P5
    want = {0: 72, 1: 120, 2: 96, 3: 272, 4: 136, 5: 40}
    for which, size in want.items():
        assert lib.ref_sizeof(which) == size
    assert lib.ref_sizeof(0) == C.sizeof(ol.RefObject) == C.sizeof(api.render_entry_3d_object)
    assert lib.ref_sizeof(1) == ol.REF_EDGE_DTYPE.itemsize == api.EDGE_INFO_DTYPE.itemsize
    assert lib.ref_sizeof(6) == C.sizeof(ol.RefLoadedBitmap) == C.sizeof(api.loaded_bitmap)
    assert lib.ref_sizeof(7) == C.sizeof(ol.RefCommands) == C.sizeof(api.game_render_commands)
    assert lib.ref_sizeof(8) == C.sizeof(ol.RefLightData) == C.sizeof(api.light_data)
    assert lib.ref_sizeof(9) == C.sizeof(ol.RefLightInfo) == C.sizeof(api.light_info)
    assert lib.ref_sizeof(10) == C.sizeof(ol.RefTransform) == C.sizeof(api.projective_transform)


@pytest.mark.parametrize("seed,n,rmin,rmax,jitter", [(1, 40_000, 1.5, 4.0, 0.5), (2, 40_000, 1.0, 8.0, 2.5),
                                                      (3, 3_000, 32.0, 96.0, 0.5), (4, 3_000, 10.0, 120.0, 2.5)])
def test_level1_walk_equals_verbatim_draw_model(seed, n, rmin, rmax, jitter):
    """Per-triangle: verbatim FillEdgeTable+DrawModel vs the port, bit-exact wherever the
    reference survives; and the port predicts exactly which triangles crash it."""
    s = sc.triangle_soup("t", seed, n, 1280, 720, rmin, rmax, jitter=jitter)
    o = ol.oracle_render(s)
    r = ol.ref_render_triangles(s)                                   # no skipping: crashes are caught
    crashed = r["status"] == -2
    assert np.array_equal(crashed, o["would_crash"].astype(bool))
    r2 = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True)
    assert np.array_equal(r2["z"].view(np.uint32), o["z"].view(np.uint32))
    assert np.array_equal(r2["color"], o["color"])


def test_reference_threaded_harness_is_exact():
    s = sc.triangle_soup("t", 9, 50_000, 1280, 720, 1.5, 12.0)
    o = ol.oracle_render(s)
    r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True, threads=5)
    assert np.array_equal(r["z"].view(np.uint32), o["z"].view(np.uint32))
    assert np.array_equal(r["color"], o["color"])


def test_edge_tables_random_objects():
    s = sc.triangle_soup("t", 21, 5_000, 1920, 1080, 1.0, 60.0, jitter=2.5)
    e_ref, n_ref = ol.ref_edge_table(s)
    e_orc, n_orc = ol.oracle_edge_table(s)
    assert n_ref == n_orc
    for f in ol.GOURAUD_FIELDS:
        assert np.array_equal(np.ascontiguousarray(e_ref[f]).view(np.uint32),
                              np.ascontiguousarray(e_orc[f]).view(np.uint32)), f

"""The oracle's whole-object mode (projekt.cpp:198-303, 542-597 replayed link by
link) against golden vectors from the verbatim call pair FillEdgeTable + DrawModel on WHOLE objects,
including the inputs on which the verbatim build dereferences a null list pointer."""
import os

import numpy as np
import pytest

import level0_cases
import oracle_lib as ol

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_level0.npz"))
BASE = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
CASES = level0_cases.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_level0_against_verbatim_golden(name):
    s, phong = CASES[name]
    o = ol.oracle_render_object(s, phong=phong)
    ref_crashed = int(GOLD[f"{name}_status"]) < 0
    assert bool(o["status"] & 2) == ref_crashed              # the oracle predicts the crash ...
    assert ol.fnv1a64_words(o["z"]) == str(GOLD[f"{name}_z_hash"])       # ... and stops where it happens
    assert ol.fnv1a64_words(o["color"]) == str(GOLD[f"{name}_color_hash"])


def test_level0_differs_from_level1_where_the_reference_mispairs_edges():
    """SURVEY.md section 0, probe P4: 419 depth / 221 colour pixels of 89 833 on C1 at 1080p."""
    from cpu_renderer_b200 import scene as sc
    m = level0_cases.MESH
    s = sc.sphere_scene(m["pos"], m["col"], m["nrm"], m["uvs"], 1920, 1080, 500.0)
    o0 = ol.oracle_render_object(s)
    o1 = ol.oracle_render(s)
    assert ol.fnv1a64_words(o0["z"]) == str(BASE["c1_1080p_level0_z_hash"])
    assert int((o0["z"].view(np.uint32) != o1["z"].view(np.uint32)).sum()) == int(BASE["c1_1080p_level01_z_diff"])
    assert int((o0["color"] != o1["color"]).sum()) == int(BASE["c1_1080p_level01_color_diff"])


def test_span_owners_follow_draw_order():
    s, _ = CASES["sphere"]
    o = ol.oracle_render_object(s, with_prim=True, prim_base=1000)
    covered = o["z"] != np.float32(s.clear_depth)
    assert o["prim"][covered].min() >= 1000 and (o["prim"][~covered] == -1).all()
    rows = np.nonzero(covered.any(axis=1))[0]
    first = [o["prim"][r][covered[r]].min() for r in rows]
    assert all(a < b for a, b in zip(first, first[1:]))      # rows are drawn top to bottom


@pytest.mark.skipif(not ol.ref_available(), reason="no verbatim reference build")
@pytest.mark.parametrize("name", ["sphere_tex_phong", "sphere_moved1", "soup_as_object3"])
def test_level0_live_against_verbatim(name):
    s, phong = CASES[name]
    o = ol.oracle_render_object(s, phong=phong)
    r = ol.ref_render_object(s, phong=phong)
    assert bool(o["status"] & 2) == (r["status"] < 0)
    assert np.array_equal(o["z"].view(np.uint32), r["z"].view(np.uint32))
    assert np.array_equal(o["color"], r["color"])

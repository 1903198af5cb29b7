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


def test_single_triangle_objects_level0_equals_level1():
    """Two independent restatements of DrawModel's walk live in the oracle: the per-triangle one (arrays,
    <= 3 edges, defined behaviour where the reference dereferences null) and the whole-object one (the
    pointer graph, link by link).  For an object that IS one triangle they must produce the same pixels
    whenever the reference survives; where it does not, level 0 stops and level 1 goes on by definition."""
    from dataclasses import replace
    from cpu_renderer_b200 import scene as sc
    soup = sc.triangle_soup("prop", 0xFEED, 3000, 400, 300, 1.0, 60.0, jitter=2.5)
    agree = stopped = 0
    for t in range(soup.triangle_count):
        one = replace(soup, positions=soup.positions[3*t:3*t + 3], colors=soup.colors[3*t:3*t + 3],
                      normals=soup.normals[3*t:3*t + 3], uvs=soup.uvs[3*t:3*t + 3])
        l0 = ol.oracle_render_object(one)
        l1 = ol.oracle_render(one)
        if l0["status"] > 0 and (l0["status"] & 2):
            stopped += 1
            assert l1["would_crash"][0] == 1, t           # both predict the same crash
            # level 0 drew a prefix of what level 1 draws
            drawn0 = l0["z"] != np.float32(one.clear_depth)
            assert np.array_equal(l0["z"][drawn0].view(np.uint32), l1["z"][drawn0].view(np.uint32)), t
            continue
        assert l1["would_crash"][0] == 0, t
        assert np.array_equal(l0["z"].view(np.uint32), l1["z"].view(np.uint32)), t
        assert np.array_equal(l0["color"], l1["color"]), t
        agree += 1
    assert agree > 2500 and stopped > 0, (agree, stopped)


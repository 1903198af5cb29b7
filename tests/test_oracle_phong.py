"""The oracle's per-pixel Phong path (projekt.cpp:450-509, 4012-4019, 551-552)
against golden vectors from the verbatim reference, and live against the verbatim build."""
import os

import numpy as np
import pytest

import kat_scenes
import oracle_lib as ol
from cpu_renderer_b200 import scene as sc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_phong.npz"))
MESH = np.load(os.path.join(os.path.dirname(__file__), "golden", "sphere_mesh.npz"))


@pytest.mark.parametrize("name", sorted(kat_scenes.all_scenes()))
def test_phong_kat_scene(name):
    s = kat_scenes.all_scenes()[name]
    o = ol.oracle_render(s, phong=True)
    assert np.array_equal(o["z"].view(np.uint32), GOLD[f"kat_{name}_z"])
    assert np.array_equal(o["color"], GOLD[f"kat_{name}_color"])


@pytest.mark.parametrize("name,kw", [
    ("soup_small", dict(seed=0xB2000002, count=30_000, width=1280, height=720, rmin=1.5, rmax=6.0)),
    ("soup_large", dict(seed=0xB2000003, count=1_500, width=1280, height=720, rmin=32.0, rmax=96.0)),
])
def test_phong_soup_hash(name, kw):
    s = sc.triangle_soup(name, **kw)
    o = ol.oracle_render(s, phong=True)
    assert ol.fnv1a64_words(o["color"]) == str(GOLD[f"{name}_color_hash"])
    assert ol.fnv1a64_words(o["z"]) == str(GOLD[f"{name}_z_hash"])
    # shading does not change coverage or depth (same edges, same z arithmetic)
    g = ol.oracle_render(s)
    assert np.array_equal(g["z"].view(np.uint32), o["z"].view(np.uint32))
    assert (g["color"] != o["color"]).sum() > 1000


def test_phong_sphere_edges_and_image():
    s = sc.sphere_scene(MESH["pos"], MESH["col"], MESH["nrm"], MESH["uvs"], 960, 540, 135.0)
    e, n = ol.oracle_edge_table(s, phong=True)
    words = np.concatenate([np.ascontiguousarray(e[f]).view(np.uint32).reshape(len(e), -1)
                            for f in ol.PHONG_FIELDS], axis=1)
    assert np.array_equal(words, GOLD["sphere_540p_edges"])
    o = ol.oracle_render(s, phong=True)
    assert np.array_equal(o["color"], GOLD["sphere_540p_color"])
    assert ol.fnv1a64_words(o["z"]) == str(GOLD["sphere_540p_z_hash"])


@pytest.mark.skipif(not ol.ref_available(), reason="no verbatim reference build")
def test_phong_live_against_verbatim():
    s = sc.triangle_soup("w", 0x5151, 8000, 800, 600, 1.0, 40.0, jitter=2.5)
    s.lights = [sc.Light(), sc.Light(P=(-4.0, 3.0, 6.0), intensity=(0.2, 0.5, 0.3, 0.1))]
    o = ol.oracle_render(s, phong=True)
    r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True, phong=True)
    assert np.array_equal(o["z"].view(np.uint32), r["z"].view(np.uint32))
    assert np.array_equal(o["color"], r["color"])

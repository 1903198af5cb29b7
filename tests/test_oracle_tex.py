"""The oracle's textured, perspective-correct path (projekt.cpp:427-446,
3919-3923, 4002-4008, 4034-4060, 4078-4089, 554-560) against golden vectors from the verbatim
reference, and live against the verbatim build.

The reference does not range-check texel coordinates.  Depth never depends on the texture, so it is
pinned to the verbatim build in every scene; colour is pinned to the verbatim build wherever every
texel coordinate stays inside the bitmap, and to the oracle's DEFINED clamp elsewhere (the golden
file stores how many pixels that concerns)."""
import os
from dataclasses import replace

import numpy as np
import pytest

import kat_scenes
import oracle_lib as ol
from cpu_renderer_b200 import scene as sc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_tex.npz"))
MESH = np.load(os.path.join(os.path.dirname(__file__), "golden", "sphere_mesh.npz"))
SOUPS = [("soup_small", dict(seed=0xB2000002, count=30_000, width=1280, height=720, rmin=1.5, rmax=6.0)),
         ("soup_large", dict(seed=0xB2000003, count=1_500, width=1280, height=720, rmin=32.0, rmax=96.0))]


def tex_soup(name, kw):
    return sc.textured(sc.triangle_soup(name, **kw), 256, 128, lo=0.4, hi=0.6)


def tex_sphere():
    s = sc.sphere_scene(MESH["pos"], MESH["col"], MESH["nrm"], MESH["uvs"], 960, 540, 135.0)
    return replace(s, texture=sc.make_texture(64, 48))


def test_texel_channels_survive_the_float_round_trip():
    """projekt.cpp:440-443 then :520-523: Round((b/255)*255) == b for every byte, so an unlit
    textured pixel is the texel word itself."""
    b = np.arange(256, dtype=np.float32)
    assert np.array_equal(np.rint((b / np.float32(255.0)) * np.float32(255.0)).astype(np.int64), np.arange(256))


@pytest.mark.parametrize("phong", [False, True])
@pytest.mark.parametrize("name", sorted(kat_scenes.all_scenes()))
def test_textured_kat_scene(name, phong):
    tag = "phong" if phong else "gouraud"
    s = sc.textured(kat_scenes.all_scenes()[name])
    o = ol.oracle_render(s, phong=phong)
    assert np.array_equal(o["z"].view(np.uint32), GOLD[f"{tag}_kat_{name}_z"])
    assert np.array_equal(o["color"], GOLD[f"{tag}_kat_{name}_color"])
    assert o["stats"]["TexelClamps"] == int(GOLD[f"{tag}_kat_{name}_clamps"])


def test_unlit_textured_pixels_are_texel_words():
    s = sc.textured(kat_scenes.all_scenes()["ties"])
    o = ol.oracle_render(s)
    covered = o["z"] != np.float32(s.clear_depth)
    assert covered.sum() > 1000
    assert np.isin(o["color"][covered], s.texture.ravel()).all()


@pytest.mark.parametrize("phong", [False, True])
@pytest.mark.parametrize("name,kw", SOUPS)
def test_textured_soup_hash(name, kw, phong):
    tag = "phong" if phong else "gouraud"
    s = tex_soup(name, kw)
    o = ol.oracle_render(s, phong=phong)
    assert o["stats"]["TexelClamps"] == 0
    assert ol.fnv1a64_words(o["color"]) == str(GOLD[f"{tag}_{name}_color_hash"])
    assert ol.fnv1a64_words(o["z"]) == str(GOLD[f"{tag}_{name}_z_hash"])
    # a texture does not change coverage or depth
    g = ol.oracle_render(replace(s, texture=None), phong=phong)
    assert np.array_equal(g["z"].view(np.uint32), o["z"].view(np.uint32))
    assert (g["color"] != o["color"]).sum() > 1000


@pytest.mark.parametrize("phong", [False, True])
def test_textured_sphere_edge_table(phong):
    tag = "phong" if phong else "gouraud"
    e, n = ol.oracle_edge_table(tex_sphere(), phong=phong)
    fields = (ol.PHONG_FIELDS if phong else ol.GOURAUD_FIELDS) + ol.TEX_FIELDS
    words = np.concatenate([np.ascontiguousarray(e[f]).view(np.uint32).reshape(len(e), -1) for f in fields], axis=1)
    assert np.array_equal(words, GOLD[f"{tag}_sphere_540p_edges"])


def test_out_of_range_uvs_are_clamped_to_the_bitmap():
    """The defined behaviour where the reference reads outside its texture: UVs far outside [0,1]
    sample the border texels, NaN samples texel (0, 0)."""
    s = sc.textured(kat_scenes.all_scenes()["ties"], 16, 8)
    uv = s.uvs.copy()
    uv[0::3] = (7.5, -3.0); uv[1::3] = (9.0, -2.0); uv[2::3] = (8.0, -4.0)      # u >> 1, v << 0
    o = ol.oracle_render(replace(s, uvs=uv))
    covered = o["z"] != np.float32(s.clear_depth)
    assert o["stats"]["TexelClamps"] == o["stats"]["Fragments"]
    assert (o["color"][covered] == s.texture[0, 15]).all()
    o = ol.oracle_render(replace(s, uvs=np.full_like(uv, np.nan)))
    assert (o["color"][covered] == s.texture[0, 0]).all()


@pytest.mark.skipif(not ol.ref_available(), reason="no verbatim reference build")
@pytest.mark.parametrize("phong", [False, True])
def test_textured_live_against_verbatim(phong):
    s = sc.textured(sc.triangle_soup("w", 0x5151, 8000, 800, 600, 1.0, 40.0, jitter=2.5), 128, 96, lo=0.4, hi=0.6)
    s.lights = [sc.Light(), sc.Light(P=(-4.0, 3.0, 6.0), intensity=(0.2, 0.5, 0.3, 0.1))]
    o = ol.oracle_render(s, with_prim=True, phong=phong)
    r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True, phong=phong)
    assert np.array_equal(o["z"].view(np.uint32), r["z"].view(np.uint32))
    if o["stats"]["TexelClamps"] == 0:
        assert np.array_equal(o["color"], r["color"])
    e_ref, n_ref = ol.ref_edge_table(s, phong=phong)
    e_orc, n_orc = ol.oracle_edge_table(s, phong=phong)
    assert n_ref == n_orc
    for f in (ol.PHONG_FIELDS if phong else ol.GOURAUD_FIELDS) + ol.TEX_FIELDS:
        a, b = np.ascontiguousarray(e_ref[f]), np.ascontiguousarray(e_orc[f])
        same = a.view(np.uint32) == b.view(np.uint32)
        if a.dtype.kind == "f":
            same |= (np.isnan(a) & np.isnan(b))             # NaN payload is not part of the contract
        assert same.all(), f

#!/usr/bin/env python
"""Generate tests/golden/*.npz from the VERBATIM reference (oracle/_ref/libprojekt_ref.so).

The reference ships no tests, golden images or known-answer vectors, so the
pins are outputs of the reference's own scalar functions -- ConstructSphere, ProjectVertex,
MergeSort, FillEdgeTable, DrawModel (projekt.cpp:2-601, 3882-4289) -- compiled unmodified behind
oracle/ref_shim.h.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Everything written here is small (< 1 MB in total) and committed; the tests that read it
(tests/test_oracle_golden.py) need neither /root/reference nor the _ref library.
"""
import os
import sys

import numpy as np
from dataclasses import replace

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from cpu_renderer_b200 import scene as sc  # noqa: E402
import kat_scenes  # noqa: E402


def edge_fields(e):
    """Pack the Gouraud-defined fields of a verbatim edge_info array as raw u32 words."""
    cols = []
    for f in ol.GOURAUD_FIELDS:
        a = np.ascontiguousarray(e[f]).view(np.uint32).reshape(len(e), -1)
        cols.append(a)
    return np.concatenate(cols, axis=1)


def main():
    lib = ol.ref()
    out = {}

    # ---- ConstructSphere (projekt.cpp:4123): the C1 mesh itself -------------------------------
    pos, col, nrm, uvs = ol.ref_sphere()
    np.savez_compressed(os.path.join(HERE, "sphere_mesh.npz"), pos=pos, col=col, nrm=nrm, uvs=uvs)

    # ---- C1: sphere as ONE object, level 0 (verbatim FillEdgeTable + DrawModel) ---------------
    for tag, (w, h, m2p) in {"c1_1080p": (1920, 1080, 500.0), "c1_540p": (960, 540, 135.0)}.items():
        s = sc.sphere_scene(pos, col, nrm, uvs, w, h, m2p)
        e, n = ol.ref_edge_table(s)
        r0 = ol.ref_render_object(s)
        assert r0["status"] == n, r0["status"]
        out[f"{tag}_edge_count"] = np.int64(n)
        out[f"{tag}_edges"] = edge_fields(e)
        out[f"{tag}_level0_color_hash"] = np.array(ol.fnv1a64_words(r0["color"]))
        out[f"{tag}_level0_z_hash"] = np.array(ol.fnv1a64_words(r0["z"]))
        out[f"{tag}_level0_covered"] = np.int64((r0["z"] != np.float32(s.clear_depth)).sum())
        # level 1 (one triangle = one object) through the verbatim call pair, crash-prone
        # triangles (predicted by the oracle) through the oracle port
        o = ol.oracle_render(s)
        r1 = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True)
        out[f"{tag}_level1_color_hash"] = np.array(ol.fnv1a64_words(r1["color"]))
        out[f"{tag}_level1_z_hash"] = np.array(ol.fnv1a64_words(r1["z"]))
        out[f"{tag}_level1_ref_crashes"] = np.int64(o["would_crash"].sum())
        out[f"{tag}_level01_z_diff"] = np.int64((r0["z"].view(np.uint32) != r1["z"].view(np.uint32)).sum())
        out[f"{tag}_level01_color_diff"] = np.int64((r0["color"] != r1["color"]).sum())

    # ---- ProjectVertex KAT (projekt.cpp:74-93), including the near-plane collapse -------------
    rng = np.random.default_rng(7)
    cam = rng.uniform(-6, 6, size=(256, 3)).astype(np.float32)
    cam[:8, 2] = np.float32([9.79, 9.8, 9.81, 10.0, 10.5, 9.799999, 9.800001, 50.0])   # around D - 0.2
    tr = ol.RefTransform(540.0, (ol.C.c_float * 2)(960.0, 540.0), 1.0, 10.0)
    prj = np.zeros_like(cam)
    for i in range(len(cam)):
        lib.ref_project_vertex(cam[i].ctypes.data_as(ol.f32p), ol.C.byref(tr), prj[i].ctypes.data_as(ol.f32p))
    out["project_in"] = cam
    out["project_out"] = prj.view(np.uint32)

    # ---- MergeSort KAT (projekt.cpp:2-72): tie order is NOT stable ----------------------------
    keys_all, perm_all = [], []
    for n in (1, 2, 3, 4, 5, 7, 8, 16, 33, 100):
        keys = rng.integers(0, 4, size=n).astype(np.int32)
        e = np.zeros(n, dtype=ol.REF_EDGE_DTYPE)
        e["YMin"] = keys
        e["YMax"] = np.arange(n)                     # identity tag
        tmp = np.zeros(n, dtype=ol.REF_EDGE_DTYPE)
        assert lib.ref_merge_sort(n, e.ctypes.data, tmp.ctypes.data) == 0
        keys_all.append(keys); perm_all.append(e["YMax"].astype(np.int32))
    out["mergesort_keys"] = np.concatenate(keys_all)
    out["mergesort_perm"] = np.concatenate(perm_all)
    out["mergesort_sizes"] = np.array([1, 2, 3, 4, 5, 7, 8, 16, 33, 100], dtype=np.int32)

    # ---- per-function KAT scenes: clip at y<0, horizontal edges, span clamps, equal-Z ties -----
    for name, s in kat_scenes.all_scenes().items():
        o = ol.oracle_render(s)
        r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True)
        out[f"kat_{name}_color"] = r["color"]
        out[f"kat_{name}_z"] = r["z"].view(np.uint32)
        out[f"kat_{name}_status"] = r["status"]
        # per-triangle edge tables straight from verbatim FillEdgeTable
        counts, words = [], []
        for t in range(s.triangle_count):
            e, n = ol.ref_edge_table(s, first_vertex=3 * t, vertex_count=3)
            counts.append(n)
            if n > 0:
                words.append(edge_fields(e))
        out[f"kat_{name}_edge_counts"] = np.array(counts, dtype=np.int32)
        out[f"kat_{name}_edges"] = np.concatenate(words) if words else np.zeros((0, 15), np.uint32)

    # ---- random soups: frame hashes of the verbatim per-triangle render -----------------------
    for name, kw in {"soup_small": dict(seed=0xB2000002, count=60_000, width=1920, height=1080, rmin=1.5, rmax=4.0),
                     "soup_large": dict(seed=0xB2000003, count=3_000, width=1920, height=1080, rmin=32.0, rmax=96.0),
                     "soup_wild": dict(seed=0x5151, count=20_000, width=800, height=600, rmin=1.0, rmax=40.0, jitter=2.5)}.items():
        s = sc.triangle_soup(name, **kw)
        o = ol.oracle_render(s)
        r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True)
        out[f"{name}_color_hash"] = np.array(ol.fnv1a64_words(r["color"]))
        out[f"{name}_z_hash"] = np.array(ol.fnv1a64_words(r["z"]))
        out[f"{name}_ref_crashes"] = np.int64(o["would_crash"].sum())
        out[f"{name}_covered"] = np.int64((r["z"] != np.float32(s.clear_depth)).sum())

    # ---- per-pixel Phong path (projekt.cpp:450-509, 4012-4019): verbatim images and hashes -----
    phong = {}
    for name, s in kat_scenes.all_scenes().items():
        o = ol.oracle_render(s, phong=True)
        r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True, phong=True)
        phong[f"kat_{name}_color"] = r["color"]
        phong[f"kat_{name}_z"] = r["z"].view(np.uint32)
    for name, kw in {"soup_small": dict(seed=0xB2000002, count=30_000, width=1280, height=720, rmin=1.5, rmax=6.0),
                     "soup_large": dict(seed=0xB2000003, count=1_500, width=1280, height=720, rmin=32.0, rmax=96.0)}.items():
        s = sc.triangle_soup(name, **kw)
        o = ol.oracle_render(s, phong=True)
        r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True, phong=True)
        phong[f"{name}_color_hash"] = np.array(ol.fnv1a64_words(r["color"]))
        phong[f"{name}_z_hash"] = np.array(ol.fnv1a64_words(r["z"]))
    s = sc.sphere_scene(pos, col, nrm, uvs, 960, 540, 135.0)
    e, n = ol.ref_edge_table(s, phong=True)
    cols = [np.ascontiguousarray(e[f]).view(np.uint32).reshape(len(e), -1) for f in ol.PHONG_FIELDS]
    phong["sphere_540p_edges"] = np.concatenate(cols, axis=1)
    o = ol.oracle_render(s, phong=True)
    r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True, phong=True)
    phong["sphere_540p_color"] = r["color"]
    phong["sphere_540p_z_hash"] = np.array(ol.fnv1a64_words(r["z"]))
    np.savez_compressed(os.path.join(HERE, "reference_vectors_phong.npz"), **phong)

    # ---- textured path (projekt.cpp:427-446, 4002-4008, 4078-4089), Gouraud and Phong ---------------
    # Depth is always the verbatim build's.  Colour is the verbatim build's wherever every texel
    # coordinate stays inside the bitmap (clamps == 0); where the reference reads outside its texture
    # (undefined) the golden colour is the oracle's defined clamp, and the number of such pixels is stored.
    tex = {}
    for phong in (False, True):
        tag = "phong" if phong else "gouraud"
        for name, s0 in kat_scenes.all_scenes().items():
            s = sc.textured(s0)
            o = ol.oracle_render(s, phong=phong)
            r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True, phong=phong)
            clamps = o["stats"]["TexelClamps"]
            assert (o["z"].view(np.uint32) == r["z"].view(np.uint32)).all(), name
            if clamps == 0:
                assert (o["color"] == r["color"]).all(), name
            tex[f"{tag}_kat_{name}_z"] = r["z"].view(np.uint32)
            tex[f"{tag}_kat_{name}_color"] = r["color"] if clamps == 0 else o["color"]
            tex[f"{tag}_kat_{name}_clamps"] = np.int64(clamps)
        for name, kw in {"soup_small": dict(seed=0xB2000002, count=30_000, width=1280, height=720, rmin=1.5, rmax=6.0),
                         "soup_large": dict(seed=0xB2000003, count=1_500, width=1280, height=720, rmin=32.0, rmax=96.0)}.items():
            # UVs in [0.4, 0.6]: spans overshoot tiny triangles by a pixel, and with a wider range the
            # steep UV gradients there extrapolate out of the bitmap (481 pixels at [0.1, 0.9])
            s = sc.textured(sc.triangle_soup(name, **kw), 256, 128, lo=0.4, hi=0.6)
            o = ol.oracle_render(s, phong=phong)
            r = ol.ref_render_triangles(s, skip=o["would_crash"], use_fallback=True, phong=phong)
            assert o["stats"]["TexelClamps"] == 0, (name, o["stats"]["TexelClamps"])
            tex[f"{tag}_{name}_color_hash"] = np.array(ol.fnv1a64_words(r["color"]))
            tex[f"{tag}_{name}_z_hash"] = np.array(ol.fnv1a64_words(r["z"]))
        # the demo sphere with the UVs ConstructSphere gives it (projekt.cpp:4174-4277), one object
        s = replace(sc.sphere_scene(pos, col, nrm, uvs, 960, 540, 135.0), texture=sc.make_texture(64, 48))
        e, n = ol.ref_edge_table(s, phong=phong)
        fields = (ol.PHONG_FIELDS if phong else ol.GOURAUD_FIELDS) + ol.TEX_FIELDS
        tex[f"{tag}_sphere_540p_edges"] = np.concatenate(
            [np.ascontiguousarray(e[f]).view(np.uint32).reshape(len(e), -1) for f in fields], axis=1)
    np.savez_compressed(os.path.join(HERE, "reference_vectors_tex.npz"), **tex)

    # ---- level 0: whole objects through the verbatim call pair --------------------
    # status >= 0: the edge count; -2: the reference dereferenced a null list pointer (the image holds
    # what it had drawn until then -- the oracle stops at the same place)
    lvl0 = {}
    cases = {}
    s540 = sc.sphere_scene(pos, col, nrm, uvs, 960, 540, 135.0)
    cases["sphere"] = (s540, False)
    cases["sphere_phong"] = (s540, True)
    cases["sphere_tex"] = (replace(s540, texture=sc.make_texture(64, 48)), False)
    cases["sphere_tex_phong"] = (replace(s540, texture=sc.make_texture(64, 48)), True)
    for k, (P, m2p) in enumerate([((0.9, 0.0, 0.0), 300.0), ((-0.8, -0.6, 0.0), 400.0), ((0.0, 0.75, 0.0), 350.0),
                                  ((0.3, -0.2, 0.5), 150.0), ((0.0, 0.0, 0.0), 60.0)]):
        cases[f"sphere_moved{k}"] = (sc.sphere_scene(pos, col, nrm, uvs, 960, 540, m2p, object_p=P), False)
    for seed, cnt in ((1, 40), (2, 200), (3, 1000)):
        cases[f"soup_as_object{seed}"] = (sc.triangle_soup("one", seed, cnt, 640, 360, 4.0, 30.0), False)
    for name, (s, phong) in cases.items():
        r = ol.ref_render_object(s, phong=phong)
        lvl0[f"{name}_status"] = np.int64(r["status"])
        lvl0[f"{name}_color_hash"] = np.array(ol.fnv1a64_words(r["color"]))
        lvl0[f"{name}_z_hash"] = np.array(ol.fnv1a64_words(r["z"]))
    np.savez_compressed(os.path.join(HERE, "reference_vectors_level0.npz"), **lvl0)

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    for k in sorted(out):
        v = out[k]
        print(k, v if v.size == 1 else v.shape)


if __name__ == "__main__":
    main()

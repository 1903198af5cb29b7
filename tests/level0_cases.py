"""Whole-object (level 0) test cases shared by the golden generator's consumers: name -> (scene, phong)."""
import os
from dataclasses import replace

import numpy as np

from cpu_renderer_b200 import scene as sc

MESH = np.load(os.path.join(os.path.dirname(__file__), "golden", "sphere_mesh.npz"))
MOVES = [((0.9, 0.0, 0.0), 300.0), ((-0.8, -0.6, 0.0), 400.0), ((0.0, 0.75, 0.0), 350.0),
         ((0.3, -0.2, 0.5), 150.0), ((0.0, 0.0, 0.0), 60.0)]


def cases():
    pos, col, nrm, uvs = MESH["pos"], MESH["col"], MESH["nrm"], MESH["uvs"]
    s540 = sc.sphere_scene(pos, col, nrm, uvs, 960, 540, 135.0)
    out = {"sphere": (s540, False), "sphere_phong": (s540, True),
           "sphere_tex": (replace(s540, texture=sc.make_texture(64, 48)), False),
           "sphere_tex_phong": (replace(s540, texture=sc.make_texture(64, 48)), True)}
    for k, (P, m2p) in enumerate(MOVES):
        out[f"sphere_moved{k}"] = (sc.sphere_scene(pos, col, nrm, uvs, 960, 540, m2p, object_p=P), False)
    for seed, cnt in ((1, 40), (2, 200), (3, 1000)):
        out[f"soup_as_object{seed}"] = (sc.triangle_soup("one", seed, cnt, 640, 360, 4.0, 30.0), False)
    return out

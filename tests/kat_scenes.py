"""Hand-built known-answer scenes for the reference's edge cases.

Each scene is a handful of triangles at 320x200 whose verbatim-reference render is stored in
tests/golden/reference_vectors.npz (tests/golden/make_golden.py).  The reference tests none of
this itself; the cases are the quirks listed in the survey:
  top clip (projekt.cpp:3993-3997, 4075-4076, 4091), edges wholly above the screen (:3968),
  exactly horizontal edges (:4066), span clamps at x<0 / x>=Width incl. "paints column 0 /
  Width-1" (:381-400), bottom clip (:192-196), equal depth -> first submitted wins (:525),
  back faces (:3943), vertices behind the near plane (:86-92), several lights (:4022-4062).
"""
from __future__ import annotations

import zlib

import numpy as np

from cpu_renderer_b200 import scene as sc

W, H = 320, 200


def _transform():
    return sc.Transform(meters_to_pixels=H / 2.0, screen_center=(W / 2.0, H / 2.0), focal_length=1.0,
                        distance_above_target=10.0)


def _unproject(sx, sy, z, tr):
    f32 = np.float32
    dist = (f32(tr.distance_above_target) - f32(z)) / f32(tr.focal_length)
    inv_m = f32(1.0) / f32(tr.meters_to_pixels)
    return [dist * ((f32(sx) - f32(tr.screen_center[0])) * inv_m),
            dist * ((f32(sy) - f32(tr.screen_center[1])) * inv_m), f32(z)]


def _orient(pts):
    """Order three screen-space points so the reference's back-face test passes."""
    (x0, y0, _), (x1, y1, _), (x2, y2, _) = pts
    cz = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0)
    return pts if cz < 0 else [pts[0], pts[2], pts[1]]


def _build(name, tris, colors=None, lights=None, object_p=(0.0, 0.0, 0.0), orient=True, camera_space=False):
    tr = _transform()
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    pos, col, nrm = [], [], []
    for i, t in enumerate(tris):
        t = _orient(list(t)) if (orient and not camera_space) else list(t)
        for (a, b, c) in t:
            pos.append([a, b, c] if camera_space else _unproject(a, b, c, tr))
        if colors is not None:
            col.extend(colors[i])
        else:
            col.extend([[*rng.uniform(0.1, 1.0, 3), 1.0] for _ in range(3)])
        for _ in range(3):
            n = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.4, 0.4), 1.0])
            nrm.append(n / np.linalg.norm(n))
    pos = np.asarray(pos, np.float32); col = np.asarray(col, np.float32); nrm = np.asarray(nrm, np.float32)
    uvs = np.zeros((len(pos), 2), np.float32)
    s = sc.Scene(name, W, H, tr, pos, col, nrm, uvs, object_p=object_p)
    if lights is not None:
        s.lights = lights
    return s


def clip_top():
    return _build("clip_top", [
        [(40, -30, 1.0), (90, 60, 0.5), (10, 50, -1.0)],        # one vertex above the screen
        [(150, -80, 2.0), (220, -10, 1.0), (160, 70, 0.0)],     # two above: an edge wholly above (:3968)
        [(250, -50, 0.0), (300, -5, 0.0), (260, -20, 0.0)],     # entirely above
        [(100, -0.4, 0.3), (140, 30.2, 0.1), (80, 25.7, 0.2)],  # starts between row -0.5 and 0
        [(200, 0.0, 0.3), (240, 40.0, 0.1), (190, 35.0, 0.2)],  # vertex exactly on y = 0
    ])


def clamp_sides():
    return _build("clamp_sides", [
        [(-60, 40, 1.0), (50, 80, 0.5), (-20, 120, -1.0)],      # crosses the left border
        [(-90, 20, 1.0), (-30, 60, 0.5), (-70, 90, 0.0)],       # wholly left: paints column 0
        [(280, 30, 1.0), (380, 70, 0.5), (300, 130, 0.0)],      # crosses the right border
        [(340, 100, 1.0), (420, 140, 0.5), (350, 180, 0.0)],    # wholly right: paints column W-1
        [(-40, 150, 2.0), (400, 160, 2.0), (160, 190, 2.5)],    # spans the full width
        [(319.2, 10, 0.0), (319.9, 30, 0.0), (318.7, 25, 0.0)],  # hugging the last column
    ])


def bottom():
    return _build("bottom", [
        [(60, 150, 1.0), (120, 260, 0.5), (20, 230, -1.0)],     # runs off the bottom (:192-196)
        [(200, 199.4, 0.0), (260, 230, 0.0), (180, 215, 0.0)],  # first row is the last screen row
        [(280, 205, 0.0), (310, 240, 0.0), (270, 230, 0.0)],    # entirely below
    ])


def horizontal():
    # exactly equal projected y needs exactly equal camera y and z
    tris = [
        [(-1.0, -0.5, 0.0), (1.0, -0.5, 0.0), (0.2, 0.7, 0.0)],     # flat top
        [(-2.5, 1.2, 1.0), (-1.5, 0.1, 1.0), (-3.5, 0.1, 1.0)],     # flat bottom
        [(2.0, 0.3, -1.0), (3.5, 0.3, -1.0), (2.7, 0.3, -1.0)],     # all three on one line: no edges
        [(1.0, -1.5, 2.0), (1.0, -1.5, 2.0), (1.0, -1.5, 2.0)],     # a point
    ]
    both = []
    for t in tris:
        both.append(t)
        both.append([t[0], t[2], t[1]])                              # the other winding
    return _build("horizontal", both, camera_space=True)


def ties():
    a = [(60, 30, 0.5), (200, 60, 0.5), (90, 170, 0.5)]
    red = [[1, 0, 0, 1]] * 3
    green = [[0, 1, 0, 1]] * 3
    blue = [[0, 0, 1, 1]] * 3
    b = [(120, 20, 1.5), (300, 90, -2.0), (150, 150, 0.0)]
    return _build("ties", [a, a, b, a, b], colors=[red, green, blue, blue, red])


def slivers():
    rng = np.random.default_rng(11)
    tris = []
    for _ in range(60):
        cx, cy = rng.uniform(20, 300), rng.uniform(20, 180)
        ang = rng.uniform(0, np.pi)
        ln = rng.uniform(2, 90)
        wd = rng.uniform(0.01, 1.5)
        dx, dy = np.cos(ang) * ln, np.sin(ang) * ln
        nx, ny = -np.sin(ang) * wd, np.cos(ang) * wd
        z = rng.uniform(-3, 3, 3)
        tris.append([(cx - dx, cy - dy, z[0]), (cx + dx, cy + dy, z[1]), (cx + nx, cy + ny, z[2])])
    for _ in range(40):                                           # sub-pixel triangles
        cx, cy = rng.uniform(5, 315), rng.uniform(5, 195)
        o = rng.uniform(-0.7, 0.7, (3, 2))
        tris.append([(cx + o[k, 0], cy + o[k, 1], rng.uniform(-1, 1)) for k in range(3)])
    return _build("slivers", tris)


def backface_and_near():
    front = [(100, 40, 0.0), (180, 90, 0.0), (90, 120, 0.0)]
    back = [front[0], front[2], front[1]]
    tris = [_orient(front), back if _orient(front) == front else front]
    s1 = _build("tmp", tris, orient=False)
    # vertices at / behind the near plane (DistanceAboveTarget - z <= 0.2): ProjectVertex -> (0,0,0)
    near = [
        [(-0.5, -0.4, 9.9), (0.6, -0.2, 5.0), (0.1, 0.7, 4.0)],
        [(0.3, 0.2, 9.8), (1.5, 0.4, 9.81), (0.8, 1.1, 2.0)],
        [(-1.0, 0.5, 11.0), (-0.2, 1.0, 3.0), (-1.4, 1.3, 3.0)],
    ]
    both = []
    for t in near:
        both.append(t); both.append([t[0], t[2], t[1]])
    s2 = _build("tmp2", both, camera_space=True)
    return sc.Scene("backface_and_near", W, H, _transform(), np.concatenate([s1.positions, s2.positions]),
                    np.concatenate([s1.colors, s2.colors]), np.concatenate([s1.normals, s2.normals]),
                    np.concatenate([s1.uvs, s2.uvs]))


def lights_and_offset():
    rng = np.random.default_rng(5)
    tris = []
    for _ in range(30):
        cx, cy, r = rng.uniform(40, 280), rng.uniform(40, 160), rng.uniform(5, 40)
        a = rng.uniform(0, 2 * np.pi)
        tris.append([(cx + r * np.cos(a + k * 2.1), cy + r * np.sin(a + k * 2.1), rng.uniform(-3, 3)) for k in range(3)])
    lights = [sc.Light(P=(5.0, 5.0, 8.0), intensity=(0.8, 0.8, 0.8, 0.0)),
              sc.Light(P=(-6.0, 2.0, 4.0), intensity=(0.3, 0.1, 0.5, 0.2)),
              sc.Light(P=(0.0, -7.0, 9.0), intensity=(0.6, 0.9, 0.2, 0.0))]
    return _build("lights_and_offset", tris, lights=lights, object_p=(0.35, -0.2, 0.5))


def all_scenes():
    return {f.__name__: f() for f in (clip_top, clamp_sides, bottom, horizontal, ties, slivers,
                                      backface_and_near, lights_and_offset)}

"""Synthetic scenes for the rasterization hot path.

The reference ships no scene other than ``ConstructSphere`` (projekt.cpp:4123-4289) and no
camera, light or resolution constants, so the survey pins them; this module is that pin.
Everything is generated on the host with numpy float32 and consumed *identically* by the
CPU oracle and the CUDA path (both read the same camera-space float arrays).

Vertex streams use the reference layout (projekt.h:2-15, projekt.cpp:3898-3934):
non-indexed, three entries per triangle -- positions v3, colours v4 (r,g,b,a), normals v3,
uvs v2.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
DRAWS_PER_TRIANGLE = 32  # fixed stride into the SplitMix64 stream, 25 used


def splitmix64(seed: int, first: int, count: int) -> np.ndarray:
    """Outputs ``first .. first+count-1`` of the SplitMix64 stream started at ``seed``.

    Output i of the sequential generator is mix(seed + (i+1)*GOLDEN), so any slice of the
    stream can be produced in parallel."""
    with np.errstate(over="ignore"):
        idx = np.arange(first + 1, first + count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def u01(bits: np.ndarray) -> np.ndarray:
    """(next() >> 40) * 2^-24 as float32."""
    return ((bits >> np.uint64(40)).astype(np.float32)) * np.float32(2.0 ** -24)


@dataclass
class Transform:
    """projective_transform (projekt.cpp:79-89)."""
    meters_to_pixels: float
    screen_center: tuple
    focal_length: float = 1.0
    distance_above_target: float = 10.0


@dataclass
class Light:
    """light_info (projekt.cpp:4026-4027)."""
    P: tuple = (5.0, 5.0, 8.0)
    intensity: tuple = (0.8, 0.8, 0.8, 0.0)


@dataclass
class Scene:
    name: str
    width: int
    height: int
    transform: Transform
    positions: np.ndarray      # [ntri*3, 3] float32 camera space
    colors: np.ndarray         # [ntri*3, 4] float32 r,g,b,a
    normals: np.ndarray        # [ntri*3, 3] float32
    uvs: np.ndarray            # [ntri*3, 2] float32 (read unconditionally by the reference)
    object_p: tuple = (0.0, 0.0, 0.0)
    ambient: tuple = (0.2, 0.2, 0.2, 1.0)
    lights: list = field(default_factory=lambda: [Light()])
    clear_color: int = 0
    clear_depth: float = -1e30
    texture: np.ndarray | None = None   # [th, tw] uint32 ARGB8: the objects' loaded_bitmap (projekt.h:13), or None

    @property
    def triangle_count(self) -> int:
        return self.positions.shape[0] // 3


def default_transform(width: int, height: int) -> Transform:
    return Transform(meters_to_pixels=height / 2.0, screen_center=(width / 2.0, height / 2.0))


def _soup_chunk(seed, first_tri, count, width, height, rmin, rmax, tr: Transform, jitter=0.5):
    f32 = np.float32
    bits = splitmix64(seed, first_tri * DRAWS_PER_TRIANGLE, count * DRAWS_PER_TRIANGLE)
    u = u01(bits).reshape(count, DRAWS_PER_TRIANGLE)
    r = f32(rmin) + u[:, 2] * f32(rmax - rmin)
    # keep every vertex on screen: centre at least r+2 away from each border
    cx = (r + f32(2.0)) + u[:, 0] * (f32(width) - f32(2.0) * (r + f32(2.0)))
    cy = (r + f32(2.0)) + u[:, 1] * (f32(height) - f32(2.0) * (r + f32(2.0)))
    theta0 = u[:, 3] * f32(2.0 * np.pi)
    pos = np.empty((count, 3, 3), dtype=f32)
    col = np.empty((count, 3, 4), dtype=f32)
    nrm = np.empty((count, 3, 3), dtype=f32)
    inv_m = f32(1.0) / f32(tr.meters_to_pixels)
    for k in range(3):
        d = u[:, 4 + 7 * k: 4 + 7 * (k + 1)]
        # decreasing angle in screen space => the reference's back-face test passes
        ang = theta0 - f32(k * 2.0 * np.pi / 3.0) + f32(jitter) * (d[:, 0] - f32(0.5))
        sx = cx + r * np.cos(ang).astype(f32)
        sy = cy + r * np.sin(ang).astype(f32)
        z = f32(-4.0) + d[:, 1] * f32(8.0)
        # UnprojectVertex (projekt.cpp:147-160): xy = ((D - z)/f) * ((screen - centre)*(1/m))
        dist = (f32(tr.distance_above_target) - z) / f32(tr.focal_length)
        pos[:, k, 0] = dist * ((sx - f32(tr.screen_center[0])) * inv_m)
        pos[:, k, 1] = dist * ((sy - f32(tr.screen_center[1])) * inv_m)
        pos[:, k, 2] = z
        col[:, k, 0] = d[:, 2] ** 3
        col[:, k, 1] = d[:, 3] ** 3
        col[:, k, 2] = d[:, 4] ** 3
        col[:, k, 3] = f32(1.0)
        nx = d[:, 5] - f32(0.5)
        ny = d[:, 6] - f32(0.5)
        inv = f32(1.0) / np.sqrt(nx * nx + ny * ny + f32(1.0)).astype(f32)
        nrm[:, k, 0] = nx * inv
        nrm[:, k, 1] = ny * inv
        nrm[:, k, 2] = inv
    return pos, col, nrm


def triangle_soup(name, seed, count, width, height, rmin, rmax, chunk=1 << 20, jitter=0.5) -> Scene:
    """Random triangle soup in screen space mapped back to camera space."""
    tr = default_transform(width, height)
    pos = np.empty((count * 3, 3), dtype=np.float32)
    col = np.empty((count * 3, 4), dtype=np.float32)
    nrm = np.empty((count * 3, 3), dtype=np.float32)
    for first in range(0, count, chunk):
        n = min(chunk, count - first)
        p, c, m = _soup_chunk(seed, first, n, width, height, rmin, rmax, tr, jitter)
        pos[first * 3:(first + n) * 3] = p.reshape(-1, 3)
        col[first * 3:(first + n) * 3] = c.reshape(-1, 4)
        nrm[first * 3:(first + n) * 3] = m.reshape(-1, 3)
    uvs = np.zeros((count * 3, 2), dtype=np.float32)
    return Scene(name, width, height, tr, pos, col, nrm, uvs)


def _libm_table(fn_name: str, angles: np.ndarray) -> np.ndarray:
    """libm's sinf / cosf on a (small) float32 array -- the functions the reference's Sin / Cos are pinned
    to.  numpy's own float32 sin/cos are SIMD kernels that may differ in the last
    bit; the mesh is an INPUT shared bit for bit with the CPU checkers, so it uses the same libm."""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    fn = getattr(libm, fn_name)
    fn.argtypes = [ctypes.c_float]
    fn.restype = ctypes.c_float
    return np.array([fn(float(a)) for a in angles], dtype=np.float32)


def construct_sphere(step_count: int = 24, radius: float = 0.5):
    """Host-side restatement of ConstructSphere (projekt.cpp:4123-4289) with a parametric
    StepCount (the reference hard-codes 24, :4129).  4*S*S - 4*S triangles.  Bit-identical to the
    verbatim function at StepCount 24 -- positions, colours, normals and UVs
    (tests/test_oracle_golden.py) -- and to the oracle's C restatement at the C5 size: the
    sines and cosines of the 3S + 2 distinct angles come from libm's sinf / cosf, the colour ramp is
    accumulated row by row (:4286), every product and sum is one binary32 operation."""
    f32 = np.float32
    S = step_count
    pi32 = f32(3.14159265359)
    inc_incl = pi32 / f32(S)                                         # :4143
    inc_azim = (f32(2.0) * pi32) / f32(S * 2)                        # :4144
    incl = np.arange(S + 1, dtype=f32) * inc_incl                    # (r32)Index*Increment
    azim = np.arange(2 * S + 1, dtype=f32) * inc_azim
    si, ci = _libm_table("sinf", incl), _libm_table("cosf", incl)
    sa, ca = _libm_table("sinf", azim), _libm_table("cosf", azim)

    def point(i, a):                                                 # [len(i), len(a), 3]
        return np.stack([si[i][:, None] * ca[a][None, :], np.broadcast_to(ci[i][:, None], (len(i), len(a))),
                         si[i][:, None] * sa[a][None, :]], -1).astype(f32)

    I, A = np.arange(S), np.arange(2 * S)
    p1, p2, p3, p4 = point(I, A), point(I + 1, A), point(I + 1, A + 1), point(I, A + 1)
    blue = ((f32(1.0) + ca[A]) / f32(2.0)).astype(f32)               # :4165
    nblue = ((f32(1.0) + ca[A + 1]) / f32(2.0)).astype(f32)
    # colour ramp: CurrentColor += ColorIncrement after every inclination row (:4286), sequentially
    up, down = np.array([1, 0, 0, 1], f32), np.array([0, 1, 0, 1], f32)
    inc = ((down - up) / f32(S)).astype(f32)                         # :4135-4141
    cur = np.empty((S, 4), f32)
    c = up.copy()
    for i in range(S):
        cur[i] = c
        c = (c + inc).astype(f32)
    nxt = (cur + inc[None, :]).astype(f32)                           # CurrentColor + ColorIncrement

    def colour(base, b):                                             # base [S,4], b [2S] -> [S, 2S, 4]
        out = np.broadcast_to(base[:, None, :], (S, 2 * S, 4)).copy()
        out = (out + f32(0.0)).astype(f32)
        out[..., 2] = (base[:, None, 2] + b[None, :]).astype(f32)
        return out

    c_cb, c_nb, c_nnb, c_cnb = colour(cur, blue), colour(nxt, blue), colour(nxt, nblue), colour(cur, nblue)

    def uv_mid(p):
        return np.stack([(p[..., 0] + f32(1.0)) / f32(2.0), (p[..., 1] + f32(1.0)) / f32(2.0)], -1).astype(f32)

    def uv_pole(p):
        return np.stack([p[..., 0], p[..., 2]], -1).astype(f32)

    half = np.full((2 * S, 2), 0.5, f32)
    north = np.broadcast_to(np.array([0, 1, 0], f32), (2 * S, 3))
    south = np.broadcast_to(np.array([0, -1, 0], f32), (2 * S, 3))
    top_p = np.stack([north, p2[0], p3[0]], 1)                       # [2S, 3, 3]   (:4156-4189)
    top_c = np.stack([c_cb[0], c_nb[0], c_nnb[0]], 1)
    top_u = np.stack([half, uv_pole(p2[0]), uv_pole(p3[0])], 1)
    bot_p = np.stack([p1[S - 1], south, p4[S - 1]], 1)               # (:4190-4223)
    bot_c = np.stack([c_cb[S - 1], c_nb[S - 1], c_nnb[S - 1]], 1)
    bot_u = np.stack([half, uv_pole(south), uv_pole(p4[S - 1])], 1)
    m = slice(1, S - 1)                                              # (:4224-4281) two per cell
    mid_p = np.stack([np.stack([p1[m], p2[m], p3[m]], 2), np.stack([p1[m], p3[m], p4[m]], 2)], 2)
    mid_c = np.stack([np.stack([c_cb[m], c_nb[m], c_nnb[m]], 2),
                      np.stack([c_cb[m], c_nnb[m], c_cnb[m]], 2)], 2)
    mid_u = uv_mid(mid_p)
    nrm = np.concatenate([top_p.reshape(-1, 3), mid_p.reshape(-1, 3), bot_p.reshape(-1, 3)]).astype(f32)
    col = np.concatenate([top_c.reshape(-1, 4), mid_c.reshape(-1, 4), bot_c.reshape(-1, 4)]).astype(f32)
    uvs = np.concatenate([top_u.reshape(-1, 2), mid_u.reshape(-1, 2), bot_u.reshape(-1, 2)]).astype(f32)
    pos = (f32(radius) * nrm).astype(f32)
    return pos, col, nrm, uvs


def sphere_scene(pos, col, nrm, uvs, width=1920, height=1080, meters_to_pixels=500.0,
                 focal=2.0, dist=3.0, light_p=(2.0, 2.0, 2.0), name="c1_sphere", object_p=(0.0, 0.0, 0.0)) -> Scene:
    """Config C1: the reference sphere as one object."""
    tr = Transform(meters_to_pixels, (width / 2.0, height / 2.0), focal, dist)
    return Scene(name, width, height, tr, np.ascontiguousarray(pos, np.float32),
                 np.ascontiguousarray(col, np.float32), np.ascontiguousarray(nrm, np.float32),
                 np.ascontiguousarray(uvs, np.float32), object_p=object_p,
                 lights=[Light(P=light_p)])


# The named configs of BASELINE.json (sizes, seeds and radii from SURVEY.md 8d).
CONFIGS = {
    "c2": dict(seed=0xB2000002, count=1_000_000, width=1920, height=1080, rmin=1.5, rmax=4.0),
    "c3": dict(seed=0xB2000003, count=50_000, width=3840, height=2160, rmin=32.0, rmax=96.0),
    "c4": dict(seed=0xB2000004, count=20_000_000, width=16384, height=16384, rmin=2.0, rmax=10.0),
}


C5_VIEWS = 256
C5_STEP_COUNT = 708            # ConstructSphere at StepCount 708: 2 002 224 triangles


def c5_view(i: int):
    """View i of config C5 -> (Object->P, DistanceAboveTarget).  The reference API has no camera
    rotation, only Object->P and the pin-hole distance (projekt.cpp:3900, 74-93), so a view is a
    (P, DistanceAboveTarget) pair on a fixed spiral."""
    import math
    a = 2.0 * math.pi * i / 32.0
    r = 0.15 + 0.45 * i / C5_VIEWS
    return (r * math.cos(a), 0.6 * r * math.sin(a), 0.0), 3.0 + 1.5 * i / C5_VIEWS


def c5_scene(mesh, view: int, width: int = 1920, height: int = 1080) -> "Scene":
    """The C5 frame of one view: the sphere mesh (construct_sphere output) at that view's P and distance."""
    P, D = c5_view(view)
    s = sphere_scene(*mesh, width, height, 500.0, name=f"c5_view{view}", object_p=P)
    s.transform.distance_above_target = D
    return s


def make_texture(width: int, height: int, seed: int = 0x7E57) -> np.ndarray:
    """A deterministic ARGB8 texture in which neighbouring texels differ in every channel (an
    off-by-one texel cannot go unnoticed) and no two texels of a 256x256 block are equal."""
    y, x = np.mgrid[0:height, 0:width].astype(np.uint32)
    h = (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77)) ^ np.uint32(seed)
    h ^= h >> np.uint32(15); h *= np.uint32(0x2C1B3C6D); h ^= h >> np.uint32(12)
    r = (x * 7 + (h & 3)) & 0xFF
    g = (y * 11 + ((h >> 2) & 3)) & 0xFF
    b = (h >> 8) & 0xFF
    a = 0x80 | ((h >> 16) & 0x7F)
    return ((a << 24) | (r << 16) | (g << 8) | b).astype(np.uint32)


def textured(scene: Scene, tex_w: int = 64, tex_h: int = 48, seed: int = 0x7E57, lo: float = 0.1,
             hi: float = 0.9) -> Scene:
    """The same scene with a texture on every object: per-vertex UVs uniform in [lo, hi]^2 (SplitMix64)
    and make_texture(tex_w, tex_h).  With the default range every texel coordinate the reference
    computes stays inside the bitmap (the reference has no range check, projekt.cpp:433-438)."""
    n = scene.positions.shape[0]
    uv = (np.float32(lo) + np.float32(hi - lo) * u01(splitmix64(seed, 0, 2 * n))).astype(np.float32).reshape(n, 2)
    return replace(scene, name=scene.name + "_tex", uvs=np.ascontiguousarray(uv), texture=make_texture(tex_w, tex_h, seed))


def make_config(name: str, scale: float = 1.0) -> Scene:
    """``scale`` < 1 shrinks the triangle count (parity tests); the bench uses scale 1."""
    cfg = dict(CONFIGS[name])
    cfg["count"] = max(1, int(round(cfg["count"] * scale)))
    return triangle_soup(name, **cfg)

"""ctypes doors into the CPU checkers (TEST INFRASTRUCTURE).

``liboracle.so``            our C restatement (oracle/raster_oracle.c) -- always available.
``_ref/libprojekt_ref.so``  the verbatim reference scalar path (oracle/Makefile `ref`), built
                            where /root/reference exists; travels to the GPU box prebuilt.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REFERENCE_DIR = "/root/reference"

f32p = C.POINTER(C.c_float)


def _np_f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


# ------------------------------------------------------------------ oracle (port) structs
class OrcTransform(C.Structure):
    _fields_ = [("MetersToPixels", C.c_float), ("ScreenCenterX", C.c_float),
                ("ScreenCenterY", C.c_float), ("FocalLength", C.c_float),
                ("DistanceAboveTarget", C.c_float)]


class OrcLight(C.Structure):
    _fields_ = [("P", C.c_float * 3), ("Intensity", C.c_float * 4)]


class OrcScene(C.Structure):
    _fields_ = [("Ambient", C.c_float * 4), ("LightCount", C.c_uint32),
                ("Lights", C.POINTER(OrcLight)), ("Transform", OrcTransform)]


class OrcEdge(C.Structure):
    _fields_ = [("YMin", C.c_int32), ("YMax", C.c_int32), ("XMin", C.c_float),
                ("Gradient", C.c_float), ("ZMin", C.c_float), ("ZGradient", C.c_float),
                ("MinColor", C.c_float * 4), ("ColorGradient", C.c_float * 4),
                ("Left", C.c_int32), ("Triangle", C.c_int32),
                ("MinNormal", C.c_float * 3), ("NormalGradient", C.c_float * 3),
                ("UMin", C.c_float), ("VMin", C.c_float), ("OneOverZMin", C.c_float),
                ("UGradient", C.c_float), ("VGradient", C.c_float), ("OneOverZGradient", C.c_float)]


class OrcTexture(C.Structure):
    _fields_ = [("Width", C.c_int32), ("Height", C.c_int32), ("Pitch", C.c_int32),
                ("Memory", C.POINTER(C.c_uint32))]


ORC_EDGE_DTYPE = np.dtype([("YMin", "<i4"), ("YMax", "<i4"), ("XMin", "<f4"), ("Gradient", "<f4"),
                           ("ZMin", "<f4"), ("ZGradient", "<f4"), ("MinColor", "<f4", 4),
                           ("ColorGradient", "<f4", 4), ("Left", "<i4"), ("Triangle", "<i4"),
                           ("MinNormal", "<f4", 3), ("NormalGradient", "<f4", 3),
                           ("UMin", "<f4"), ("VMin", "<f4"), ("OneOverZMin", "<f4"),
                           ("UGradient", "<f4"), ("VGradient", "<f4"), ("OneOverZGradient", "<f4")])


class OrcTarget(C.Structure):
    _fields_ = [("Width", C.c_int32), ("Height", C.c_int32), ("Pitch", C.c_int32),
                ("Color", C.POINTER(C.c_uint32)), ("Z", f32p), ("ZStride", C.c_uint32),
                ("Prim", C.POINTER(C.c_int32))]


class OrcStats(C.Structure):
    _fields_ = [("Triangles", C.c_uint64), ("Visible", C.c_uint64), ("SpanRows", C.c_uint64),
                ("Fragments", C.c_uint64), ("DepthPasses", C.c_uint64),
                ("RefWouldCrash", C.c_uint64), ("TexelClamps", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class OrcFallbackCtx(C.Structure):
    _fields_ = [("Pos", f32p), ("Col", f32p), ("Nrm", f32p), ("P", C.c_float * 3),
                ("Scene", C.POINTER(OrcScene)), ("Phong", C.c_int32),
                ("UV", f32p), ("Texture", C.POINTER(OrcTexture))]


# ------------------------------------------------------------------ reference structs
# (layouts of oracle/ref_shim.h + /root/reference/projekt.h:2-37; sizes are asserted against
#  ref_sizeof() in tests/test_oracle_vs_ref.py)
class RefLoadedBitmap(C.Structure):
    _fields_ = [("Width", C.c_int32), ("Height", C.c_int32), ("Pitch", C.c_int32),
                ("Memory", C.c_void_p)]


class RefTransform(C.Structure):
    _fields_ = [("MetersToPixels", C.c_float), ("ScreenCenter", C.c_float * 2),
                ("FocalLength", C.c_float), ("DistanceAboveTarget", C.c_float)]


class RefLightInfo(C.Structure):
    _fields_ = [("P", C.c_float * 3), ("Intensity", C.c_float * 4)]


class RefLightData(C.Structure):
    _fields_ = [("AmbientIntensity", C.c_float * 4), ("LightCount", C.c_uint32),
                ("Lights", C.POINTER(RefLightInfo))]


class RefCommands(C.Structure):
    _fields_ = [("Width", C.c_uint32), ("ZBuffer", f32p), ("ZMask", C.c_void_p),
                ("LightData", RefLightData), ("Transform", RefTransform),
                ("ThreadMemory", C.c_void_p), ("ThreadMemorySize", C.c_uint32),
                ("ThreadMemorySizeUsed", C.c_uint32), ("SortMemory", C.c_void_p)]


class RefObject(C.Structure):
    _fields_ = [("P", C.c_float * 3), ("VertexCount", C.c_uint32), ("Optimized", C.c_int32),
                ("PhongShading", C.c_int32), ("VertexData", C.c_void_p), ("ColorData", C.c_void_p),
                ("NormalData", C.c_void_p), ("UVData", C.c_void_p), ("EdgeMemory", C.c_void_p),
                ("Bitmap", C.c_void_p)]


# edge_info, projekt.h:17-37 (120 bytes; offsets from SURVEY.md 8a row a5)
REF_EDGE_DTYPE = np.dtype({
    "names": ["YMax", "XMin", "ZMin", "OneOverZMin", "Gradient", "ZGradient", "OneOverZGradient",
              "YMin", "UMin", "VMin", "UGradient", "VGradient", "Left", "MinColor",
              "ColorGradient", "MinNormal", "NormalGradient", "Next"],
    "formats": ["<i4", "<f4", "<f4", "<f4", "<f4", "<f4", "<f4", "<i4", "<f4", "<f4", "<f4", "<f4",
                "<i4", ("<f4", 4), ("<f4", 4), ("<f4", 3), ("<f4", 3), "<u8"],
    "offsets": [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 68, 84, 96, 112],
    "itemsize": 120})

# The fields FillEdgeTable defines in Gouraud mode (SURVEY.md 8b "Ownership").
GOURAUD_FIELDS = ["YMin", "YMax", "XMin", "Gradient", "ZMin", "ZGradient", "MinColor",
                  "ColorGradient", "Left"]
# Phong mode additionally defines the normals (projekt.cpp:4017, 4104-4109)
PHONG_FIELDS = GOURAUD_FIELDS + ["MinNormal", "NormalGradient"]
# an object with a Bitmap additionally defines u/z, v/z, 1/z (projekt.cpp:4002-4004, 4078-4089)
TEX_FIELDS = ["UMin", "VMin", "OneOverZMin", "UGradient", "VGradient", "OneOverZGradient"]

REF_FALLBACK_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p)

_oracle = None
_ref = None


def build_oracle(force=False):
    """Compile liboracle.so (and _ref when the reference tree is present)."""
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src_newer = (not os.path.exists(so)) or any(
        os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(so)
        for f in ("raster_oracle.c", "raster_oracle.h"))
    if force or src_newer:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir(REFERENCE_DIR):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(os.path.join(ORACLE_DIR, "liboracle.so"))
        lib.orc_project_vertex.argtypes = [f32p, C.POINTER(OrcTransform), f32p]
        lib.orc_fill_edge_table.argtypes = [f32p, f32p, f32p, C.c_uint32, f32p, C.POINTER(OrcScene),
                                            C.c_void_p, C.c_void_p]
        lib.orc_fill_edge_table.restype = C.c_int32
        lib.orc_merge_sort.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
        lib.orc_draw_triangle.argtypes = [C.c_void_p, C.c_uint32, C.c_int32, C.POINTER(OrcTarget),
                                          C.POINTER(OrcStats)]
        lib.orc_draw_triangle.restype = C.c_int32
        lib.orc_render_triangles.argtypes = [f32p, f32p, f32p, C.c_uint32, f32p, C.POINTER(OrcScene),
                                             C.POINTER(OrcTarget), C.c_int32, C.c_void_p,
                                             C.POINTER(OrcStats)]
        lib.orc_render_triangles.restype = C.c_int32
        lib.orc_render_triangles_ex.argtypes = [f32p, f32p, f32p, C.c_uint32, f32p, C.POINTER(OrcScene),
                                                C.c_int32, C.POINTER(OrcTarget), C.c_int32, C.c_void_p,
                                                C.POINTER(OrcStats)]
        lib.orc_render_triangles_ex.restype = C.c_int32
        lib.orc_fill_edge_table_ex.argtypes = [f32p, f32p, f32p, C.c_uint32, f32p, C.POINTER(OrcScene),
                                               C.c_int32, C.c_void_p, C.c_void_p]
        lib.orc_fill_edge_table_ex.restype = C.c_int32
        lib.orc_fill_edge_table_tex.argtypes = [f32p, f32p, f32p, f32p, C.c_uint32, f32p, C.POINTER(OrcScene),
                                                C.c_int32, C.c_void_p, C.c_void_p]
        lib.orc_fill_edge_table_tex.restype = C.c_int32
        lib.orc_render_triangles_tex.argtypes = [f32p, f32p, f32p, f32p, C.c_uint32, f32p, C.POINTER(OrcScene),
                                                 C.c_int32, C.POINTER(OrcTexture), C.POINTER(OrcTarget),
                                                 C.c_int32, C.c_void_p, C.POINTER(OrcStats)]
        lib.orc_render_triangles_tex.restype = C.c_int32
        lib.orc_render_object.argtypes = [f32p, f32p, f32p, f32p, C.c_uint32, f32p, C.POINTER(OrcScene), C.c_int32,
                                          C.POINTER(OrcTexture), C.POINTER(OrcTarget), C.c_int32, C.POINTER(OrcStats)]
        lib.orc_render_object.restype = C.c_int32
        lib.orc_render_triangles_mt.argtypes = [f32p, f32p, f32p, C.c_uint32, f32p,
                                                C.POINTER(OrcScene), C.POINTER(OrcTarget),
                                                C.c_uint32, C.POINTER(OrcStats)]
        lib.orc_render_triangles_mt.restype = C.c_int32
        lib.orc_construct_sphere.argtypes = [C.c_uint32, f32p, f32p, f32p, f32p]
        lib.orc_construct_sphere.restype = C.c_uint32
        _oracle = lib
    return _oracle


def ref_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libprojekt_ref.so")) or \
        os.path.isdir(REFERENCE_DIR)


def ref():
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libprojekt_ref.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.ref_sizeof.argtypes = [C.c_uint32]
        lib.ref_sizeof.restype = C.c_uint32
        lib.ref_construct_sphere.argtypes = [f32p, f32p, f32p, f32p]
        lib.ref_construct_sphere.restype = C.c_uint32
        lib.ref_project_vertex.argtypes = [f32p, C.POINTER(RefTransform), f32p]
        lib.ref_merge_sort.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
        lib.ref_merge_sort.restype = C.c_int32
        lib.ref_fill_edge_table.argtypes = [C.POINTER(RefObject), C.POINTER(RefCommands), C.c_int32]
        lib.ref_fill_edge_table.restype = C.c_int32
        lib.ref_draw_model.argtypes = [C.POINTER(RefLoadedBitmap), C.c_void_p, C.c_uint32,
                                       C.POINTER(RefCommands), C.c_void_p, C.c_int32]
        lib.ref_draw_model.restype = C.c_int32
        lib.ref_render_object.argtypes = [C.POINTER(RefObject), C.POINTER(RefCommands),
                                          C.POINTER(RefLoadedBitmap)]
        lib.ref_render_object.restype = C.c_int32
        lib.ref_render_triangles.argtypes = [f32p, f32p, f32p, f32p, C.c_uint32, f32p,
                                             C.POINTER(RefCommands), C.POINTER(RefLoadedBitmap),
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        lib.ref_set_texture.argtypes = [C.c_void_p]
        lib.ref_set_texture.restype = None
        lib.ref_render_triangles_mt.argtypes = [f32p, f32p, f32p, f32p, C.c_uint32, f32p,
                                                C.POINTER(RefCommands), C.POINTER(RefLoadedBitmap),
                                                C.POINTER(f32p), C.c_uint32, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_int32]
        _ref = lib
    return _ref


# ------------------------------------------------------------------ scene marshalling
class OracleScene:
    """Keeps the ctypes views of a cpu_renderer_b200.scene.Scene alive."""

    def __init__(self, scene):
        self.scene = scene
        self.pos, self.pos_p = _np_f32(scene.positions)
        self.col, self.col_p = _np_f32(scene.colors)
        self.nrm, self.nrm_p = _np_f32(scene.normals)
        self.uvs, self.uvs_p = _np_f32(scene.uvs)
        self.P = (C.c_float * 3)(*scene.object_p)
        self.tex = self.orc_tex = self.ref_tex = None
        if getattr(scene, "texture", None) is not None:
            self.tex = np.ascontiguousarray(scene.texture, dtype=np.uint32)
            th, tw = self.tex.shape
            self.orc_tex = OrcTexture(tw, th, self.tex.strides[0], self.tex.ctypes.data_as(C.POINTER(C.c_uint32)))
            self.ref_tex = RefLoadedBitmap(tw, th, self.tex.strides[0], self.tex.ctypes.data)
        n = len(scene.lights)
        self.orc_lights = (OrcLight * max(n, 1))()
        self.ref_lights = (RefLightInfo * max(n, 1))()
        for i, l in enumerate(scene.lights):
            self.orc_lights[i].P[:] = l.P
            self.orc_lights[i].Intensity[:] = l.intensity
            self.ref_lights[i].P[:] = l.P
            self.ref_lights[i].Intensity[:] = l.intensity
        t = scene.transform
        self.orc = OrcScene()
        self.orc.Ambient[:] = scene.ambient
        self.orc.LightCount = n
        self.orc.Lights = C.cast(self.orc_lights, C.POINTER(OrcLight))
        self.orc.Transform = OrcTransform(t.meters_to_pixels, t.screen_center[0], t.screen_center[1],
                                          t.focal_length, t.distance_above_target)

    def ref_commands(self, zbuf: np.ndarray, sort_mem: np.ndarray | None = None) -> RefCommands:
        t = self.scene.transform
        cmd = RefCommands()
        cmd.Width = zbuf.shape[1]
        cmd.ZBuffer = zbuf.ctypes.data_as(f32p)
        cmd.LightData.AmbientIntensity[:] = self.scene.ambient
        cmd.LightData.LightCount = len(self.scene.lights)
        cmd.LightData.Lights = C.cast(self.ref_lights, C.POINTER(RefLightInfo))
        cmd.Transform.MetersToPixels = t.meters_to_pixels
        cmd.Transform.ScreenCenter[:] = t.screen_center
        cmd.Transform.FocalLength = t.focal_length
        cmd.Transform.DistanceAboveTarget = t.distance_above_target
        if sort_mem is not None:
            cmd.SortMemory = sort_mem.ctypes.data
        return cmd


def new_targets(scene, with_prim=False):
    color = np.full((scene.height, scene.width), scene.clear_color, dtype=np.uint32)
    z = np.full((scene.height, scene.width), scene.clear_depth, dtype=np.float32)
    prim = np.full((scene.height, scene.width), -1, dtype=np.int32) if with_prim else None
    return color, z, prim


def _orc_target(color, z, prim):
    t = OrcTarget()
    t.Height, t.Width = color.shape
    t.Pitch = color.strides[0]
    t.Color = color.ctypes.data_as(C.POINTER(C.c_uint32))
    t.Z = z.ctypes.data_as(f32p)
    t.ZStride = z.strides[0] // 4
    t.Prim = prim.ctypes.data_as(C.POINTER(C.c_int32)) if prim is not None else None
    return t


def oracle_render(scene, with_prim=False, threads=1, targets=None, prim_base=0, phong=False, compat=0):
    """Level-1 (one triangle = one object) render.  Returns dict(color, z, prim, stats, would_crash).
    ``compat``: ORC_COMPAT_* switches of the span fill (= the B200R_AVX_* flag values)."""
    lib = oracle()
    lib.orc_set_compat(int(compat))
    try:
        return _oracle_render(lib, scene, with_prim, threads, targets, prim_base, phong)
    finally:
        lib.orc_set_compat(0)


def _oracle_render(lib, scene, with_prim, threads, targets, prim_base, phong):
    s = OracleScene(scene)
    color, z, prim = targets if targets is not None else new_targets(scene, with_prim)
    t = _orc_target(color, z, prim)
    stats = OrcStats()
    n = scene.triangle_count
    crash = np.zeros(n, dtype=np.uint8)
    if s.orc_tex is not None:
        rc = lib.orc_render_triangles_tex(s.pos_p, s.col_p, s.nrm_p, s.uvs_p, n, s.P, C.byref(s.orc),
                                          1 if phong else 0, C.byref(s.orc_tex), C.byref(t), prim_base,
                                          crash.ctypes.data, C.byref(stats))
    elif phong:
        rc = lib.orc_render_triangles_ex(s.pos_p, s.col_p, s.nrm_p, n, s.P, C.byref(s.orc), 1, C.byref(t),
                                         prim_base, crash.ctypes.data, C.byref(stats))
    elif threads > 1:
        rc = lib.orc_render_triangles_mt(s.pos_p, s.col_p, s.nrm_p, n, s.P, C.byref(s.orc),
                                         C.byref(t), threads, C.byref(stats))
    else:
        rc = lib.orc_render_triangles(s.pos_p, s.col_p, s.nrm_p, n, s.P, C.byref(s.orc), C.byref(t),
                                      prim_base, crash.ctypes.data, C.byref(stats))
    assert rc == 0, rc
    return dict(color=color, z=z, prim=prim, stats=stats.as_dict(), would_crash=crash)


def oracle_render_object(scene, with_prim=False, targets=None, prim_base=0, phong=False):
    """Level 0: the whole scene as ONE object (orc_render_object).  status: bit 0 drew, bit 1 the
    verbatim reference would crash (the object stops drawing there); negative: no lights."""
    lib = oracle()
    s = OracleScene(scene)
    color, z, prim = targets if targets is not None else new_targets(scene, with_prim)
    t = _orc_target(color, z, prim)
    stats = OrcStats()
    rc = lib.orc_render_object(s.pos_p, s.col_p, s.nrm_p, s.uvs_p if s.orc_tex is not None else None,
                               scene.positions.shape[0], s.P, C.byref(s.orc), 1 if phong else 0,
                               C.byref(s.orc_tex) if s.orc_tex is not None else None, C.byref(t), prim_base,
                               C.byref(stats))
    return dict(color=color, z=z, prim=prim, stats=stats.as_dict(), status=rc)


def oracle_edge_table(scene, first_vertex=0, vertex_count=None, phong=False):
    """orc_fill_edge_table over (a slice of) the scene's vertex arrays as ONE object."""
    lib = oracle()
    s = OracleScene(scene)
    if vertex_count is None:
        vertex_count = scene.positions.shape[0] - first_vertex
    edges = np.zeros(max(vertex_count, 1), dtype=ORC_EDGE_DTYPE)
    tmp = np.zeros(max(vertex_count, 1), dtype=ORC_EDGE_DTYPE)
    off = first_vertex
    uv_p = s.uvs[off:].ctypes.data_as(f32p) if s.orc_tex is not None else None
    n = lib.orc_fill_edge_table_tex(s.pos[off:].ctypes.data_as(f32p), s.col[off:].ctypes.data_as(f32p),
                                    s.nrm[off:].ctypes.data_as(f32p), uv_p, vertex_count, s.P,
                                    C.byref(s.orc), 1 if phong else 0, edges.ctypes.data, tmp.ctypes.data)
    return edges[:max(n, 0)].copy(), n


def ref_edge_table(scene, first_vertex=0, vertex_count=None, phong=False):
    """Verbatim FillEdgeTable (projekt.cpp:3882) over (a slice of) the scene as ONE object."""
    lib = ref()
    s = OracleScene(scene)
    if vertex_count is None:
        vertex_count = scene.positions.shape[0] - first_vertex
    # poison so that fields the reference leaves unwritten are recognisable
    edges = np.frombuffer(bytearray(b"\xcd" * (120 * max(vertex_count, 1))), dtype=REF_EDGE_DTYPE)
    sort = np.zeros(max(vertex_count, 1), dtype=REF_EDGE_DTYPE)
    zdummy = np.zeros((1, 1), dtype=np.float32)
    cmd = s.ref_commands(zdummy, sort)
    obj = RefObject()
    obj.P[:] = scene.object_p
    obj.VertexCount = vertex_count
    obj.VertexData = s.pos.ctypes.data + first_vertex * 12
    obj.ColorData = s.col.ctypes.data + first_vertex * 16
    obj.NormalData = s.nrm.ctypes.data + first_vertex * 12
    obj.UVData = s.uvs.ctypes.data + first_vertex * 8
    obj.EdgeMemory = edges.ctypes.data
    obj.PhongShading = 1 if phong else 0
    if s.ref_tex is not None:
        obj.Bitmap = C.addressof(s.ref_tex)
    n = lib.ref_fill_edge_table(C.byref(obj), C.byref(cmd), 1 if phong else 0)
    return edges[:max(n, 0)].copy(), n


def ref_render_object(scene, targets=None, phong=False):
    """Level 0: the whole scene as ONE object through verbatim FillEdgeTable + DrawModel."""
    lib = ref()
    s = OracleScene(scene)
    color, z, _ = targets if targets is not None else new_targets(scene)
    nv = scene.positions.shape[0]
    edges = np.zeros(nv, dtype=REF_EDGE_DTYPE)
    sort = np.zeros(nv, dtype=REF_EDGE_DTYPE)
    cmd = s.ref_commands(z, sort)
    bmp = RefLoadedBitmap(scene.width, scene.height, color.strides[0], color.ctypes.data)
    obj = RefObject()
    obj.P[:] = scene.object_p
    obj.VertexCount = nv
    obj.VertexData, obj.ColorData = s.pos.ctypes.data, s.col.ctypes.data
    obj.NormalData, obj.UVData = s.nrm.ctypes.data, s.uvs.ctypes.data
    obj.EdgeMemory = edges.ctypes.data
    obj.PhongShading = 1 if phong else 0
    if s.ref_tex is not None:
        obj.Bitmap = C.addressof(s.ref_tex)
    rc = lib.ref_render_object(C.byref(obj), C.byref(cmd), C.byref(bmp))
    return dict(color=color, z=z, status=rc)


def ref_render_triangles(scene, skip=None, use_fallback=False, threads=1, targets=None, phong=False):
    """One triangle = one object through the verbatim call pair.  ``skip`` marks triangles the
    reference would crash on; with ``use_fallback`` they are drawn by the oracle port."""
    lib = ref()
    s = OracleScene(scene)
    color, z, _ = targets if targets is not None else new_targets(scene)
    n = scene.triangle_count
    status = np.zeros(n, dtype=np.int32)
    cmd = s.ref_commands(z)
    bmp = RefLoadedBitmap(scene.width, scene.height, color.strides[0], color.ctypes.data)
    skip_p = skip.ctypes.data if skip is not None else None
    fb, user = None, None
    ctx = None
    if use_fallback:
        ctx = OrcFallbackCtx(s.pos_p, s.col_p, s.nrm_p, s.P, C.pointer(s.orc), 1 if phong else 0,
                             s.uvs_p if s.orc_tex is not None else None,
                             C.pointer(s.orc_tex) if s.orc_tex is not None else None)
        fb = C.cast(oracle().orc_ref_fallback, C.c_void_p)
        user = C.cast(C.pointer(ctx), C.c_void_p)
    lib.ref_set_texture(C.addressof(s.ref_tex) if s.ref_tex is not None else None)
    if threads <= 1:
        lib.ref_render_triangles(s.pos_p, s.col_p, s.nrm_p, s.uvs_p, n, s.P, C.byref(cmd),
                                 C.byref(bmp), skip_p, status.ctypes.data, fb, user, 1 if phong else 0)
    else:
        colors = [color] + [color.copy() for _ in range(threads - 1)]
        zs = [z] + [z.copy() for _ in range(threads - 1)]
        bmps = (RefLoadedBitmap * threads)(*[
            RefLoadedBitmap(scene.width, scene.height, c.strides[0], c.ctypes.data) for c in colors])
        zptrs = (f32p * threads)(*[zz.ctypes.data_as(f32p) for zz in zs])
        lib.ref_render_triangles_mt(s.pos_p, s.col_p, s.nrm_p, s.uvs_p, n, s.P, C.byref(cmd), bmps,
                                    zptrs, threads, skip_p, fb, user, 1 if phong else 0)
    lib.ref_set_texture(None)
    return dict(color=color, z=z, status=status)


# ------------------------------------------------------------------ the reference's AVX thread-pool path
_avx = None


def avx_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libprojekt_avx.so")) or os.path.isdir(REFERENCE_DIR)


def avx():
    global _avx
    if _avx is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libprojekt_avx.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.avx_render_objects.argtypes = [C.POINTER(RefObject), C.c_uint32, C.POINTER(RefCommands),
                                           C.POINTER(RefLoadedBitmap), C.c_uint32, C.POINTER(C.c_double)]
        lib.avx_render_objects.restype = C.c_int32
        lib.avx_sizeof_row_header.restype = C.c_uint32
        _avx = lib
    return _avx


def _aligned(shape, dtype, align=64):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


class AvxFrame:
    """Everything one call of the reference's multithreaded AVX path needs:
    DrawModelOptimizedLines + FillLinesOptimized (projekt.cpp:3362-3613, 629-1490) for a list of objects that
    share one mesh and differ in Object->P.  That path handles textured + Phong objects only, loads depth with
    aligned 8-float loads (rows must be 32-byte aligned: the width is padded to 8) and fetches texels at
    trunc(u*Width) without a range check (UVs are clamped to [0, 1] and the bitmap carries one texel of padding)."""

    def __init__(self, scene, object_ps, threads):
        assert scene.texture is not None
        self.scene, self.threads = scene, int(threads)
        self.s = OracleScene(scene)
        self.uv = np.ascontiguousarray(np.clip(scene.uvs, 0.0, 1.0), np.float32)
        th, tw = scene.texture.shape
        self.tex = np.zeros((th + 1, tw + 1), np.uint32)
        self.tex[:th, :tw] = scene.texture
        self.tex[th, :tw] = scene.texture[th - 1]; self.tex[:th, tw] = scene.texture[:, tw - 1]
        self.tex_bmp = RefLoadedBitmap(tw, th, self.tex.strides[0], self.tex.ctypes.data)
        self.wpad = (scene.width + 7) // 8 * 8
        self.color = _aligned((scene.height, self.wpad), np.uint32)
        self.z = _aligned((scene.height, self.wpad), np.float32)
        self.zmask = np.zeros(scene.height * self.wpad // 8 + 64, np.uint8)
        nv = scene.positions.shape[0]
        self.edges = [np.zeros(nv, dtype=REF_EDGE_DTYPE) for _ in object_ps]
        self.sort = np.zeros(nv, dtype=REF_EDGE_DTYPE)
        self.arena = np.zeros((256 << 20) if len(object_ps) > 4 else (32 << 20), np.uint8)   # row-work arena (:2322)
        self.cmd = self.s.ref_commands(self.z, self.sort)
        self.cmd.ZMask = self.zmask.ctypes.data
        self.cmd.ThreadMemory = self.arena.ctypes.data
        self.cmd.ThreadMemorySize = self.arena.nbytes
        self.bmp = RefLoadedBitmap(scene.width, scene.height, self.color.strides[0], self.color.ctypes.data)
        self.objs = (RefObject * len(object_ps))()
        for o, P, e in zip(self.objs, object_ps, self.edges):
            o.P[:] = P
            o.VertexCount = nv
            o.Optimized, o.PhongShading = 1, 1
            o.VertexData, o.ColorData = self.s.pos.ctypes.data, self.s.col.ctypes.data
            o.NormalData, o.UVData = self.s.nrm.ctypes.data, self.uv.ctypes.data
            o.EdgeMemory = e.ctypes.data
            o.Bitmap = C.addressof(self.tex_bmp)
        self.seconds = (C.c_double * 2)()

    def clear(self):
        self.color.fill(self.scene.clear_color); self.z.fill(self.scene.clear_depth)

    def render(self):
        """-> (objects drawn or -2, seconds on the submitting thread, seconds in total)"""
        rc = avx().avx_render_objects(self.objs, len(self.objs), C.byref(self.cmd), C.byref(self.bmp), self.threads,
                                      self.seconds)
        return rc, self.seconds[0], self.seconds[1]

    def covered(self):
        return int((self.z[:, :self.scene.width] != np.float32(self.scene.clear_depth)).sum())


def ref_sphere():
    """Verbatim ConstructSphere (projekt.cpp:4123): 6 624 vertices."""
    lib = ref()
    pos = np.zeros((8192, 3), np.float32); col = np.zeros((8192, 4), np.float32)
    nrm = np.zeros((8192, 3), np.float32); uvs = np.zeros((8192, 2), np.float32)
    n = lib.ref_construct_sphere(pos.ctypes.data_as(f32p), col.ctypes.data_as(f32p),
                                 nrm.ctypes.data_as(f32p), uvs.ctypes.data_as(f32p))
    return pos[:n].copy(), col[:n].copy(), nrm[:n].copy(), uvs[:n].copy()


def oracle_sphere(step_count: int):
    """The oracle's C restatement of ConstructSphere with a StepCount parameter (projekt.cpp:4123-4289)."""
    lib = oracle()
    nv = 3 * (4 * step_count * step_count - 4 * step_count)
    pos = np.zeros((nv, 3), np.float32); col = np.zeros((nv, 4), np.float32)
    nrm = np.zeros((nv, 3), np.float32); uvs = np.zeros((nv, 2), np.float32)
    n = lib.orc_construct_sphere(step_count, pos.ctypes.data_as(f32p), col.ctypes.data_as(f32p),
                                 nrm.ctypes.data_as(f32p), uvs.ctypes.data_as(f32p))
    assert n == nv, (n, nv)
    return pos, col, nrm, uvs


def fnv1a64_words(a: np.ndarray) -> str:
    """Word-wise FNV-1a-64 over a u32 array."""
    h = 0xcbf29ce484222325
    for w in np.ascontiguousarray(a).view(np.uint32).ravel().tolist():
        h = ((h ^ w) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"

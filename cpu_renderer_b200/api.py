"""Python host side of the C ABI in include/b200_raster.h (libb200raster.so).

This module is plumbing only: it mirrors the reference's structs (projekt.h:2-37 plus the
layouts pinned in SURVEY.md Appendix A) as ctypes structures and forwards to the CUDA library.
There is no CPU implementation behind it -- if the library or a B200 is missing, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200R_LIB: another build of the same library (kernel experiments: variants compiled with other -D flags)
LIB_PATH = os.environ.get("B200R_LIB") or os.path.join(_HERE, "libb200raster.so")

OK, E_INVALID, E_CUDA, E_UNSUPPORTED, E_NOMEM, E_NO_DEVICE = 0, -1, -2, -3, -4, -5
WHOLE_OBJECT_AEL = 1
DEFER_VERDICT = 2
AVX_RIGHT_END_EXCLUSIVE = 4      # spans cover [MinX, MaxX) (projekt.cpp:782-794)
AVX_DEPTH_GE = 8                 # depth test >= (projekt.cpp:3205): last submitted wins ties
MESH_PHONG = 1

EXPORTS = [
    "b200r_create", "b200r_destroy", "b200r_last_error", "b200r_set_stream", "b200r_sync",
    "b200r_set_tile", "b200r_render_objects", "b200r_fill_edge_table", "b200r_render_device",
    "b200r_clear_device", "b200r_get_stats", "b200r_set_profiling", "b200r_get_stage_ms",
    "b200r_set_gather_target", "b200r_peer_alloc", "b200r_peer_open", "b200r_peer_release",
]
STAGES = ("setup_kernel", "tile_scan_kernel", "scatter_kernel", "raster_kernel")


class v2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class v3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class v4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class loaded_bitmap(C.Structure):
    _fields_ = [("Width", C.c_int32), ("Height", C.c_int32), ("Pitch", C.c_int32),
                ("Memory", C.c_void_p)]


class projective_transform(C.Structure):
    _fields_ = [("MetersToPixels", C.c_float), ("ScreenCenter", v2), ("FocalLength", C.c_float),
                ("DistanceAboveTarget", C.c_float)]


class light_info(C.Structure):
    _fields_ = [("P", v3), ("Intensity", v4)]


class light_data(C.Structure):
    _fields_ = [("AmbientIntensity", v4), ("LightCount", C.c_uint32),
                ("Lights", C.POINTER(light_info))]


class game_render_commands(C.Structure):
    _fields_ = [("Width", C.c_uint32), ("ZBuffer", C.c_void_p), ("ZMask", C.c_void_p),
                ("LightData", light_data), ("Transform", projective_transform),
                ("ThreadMemory", C.c_void_p), ("ThreadMemorySize", C.c_uint32),
                ("ThreadMemorySizeUsed", C.c_uint32), ("SortMemory", C.c_void_p)]


class render_entry_3d_object(C.Structure):
    _fields_ = [("P", v3), ("VertexCount", C.c_uint32), ("Optimized", C.c_int32),
                ("PhongShading", C.c_int32), ("VertexData", C.c_void_p), ("ColorData", C.c_void_p),
                ("NormalData", C.c_void_p), ("UVData", C.c_void_p), ("EdgeMemory", C.c_void_p),
                ("Bitmap", C.c_void_p)]


class device_texture(C.Structure):
    _fields_ = [("Memory", C.c_void_p), ("Width", C.c_int32), ("Height", C.c_int32), ("Pitch", C.c_int32)]


class device_mesh(C.Structure):
    _fields_ = [("Positions", C.c_void_p), ("Colors", C.c_void_p), ("Normals", C.c_void_p),
                ("TriangleCount", C.c_uint32), ("P", v3), ("Flags", C.c_uint32),
                ("UVs", C.c_void_p), ("Texture", C.POINTER(device_texture))]


class device_target(C.Structure):
    _fields_ = [("Color", C.c_void_p), ("Depth", C.c_void_p), ("Width", C.c_int32),
                ("Height", C.c_int32), ("ColorPitch", C.c_int32), ("DepthStride", C.c_int32),
                ("BandFirstRow", C.c_int32), ("BandRows", C.c_int32)]


class frame_stats(C.Structure):
    _fields_ = [("Triangles", C.c_uint64), ("Binned", C.c_uint64), ("Segments", C.c_uint64),
                ("Spans", C.c_uint64), ("AliasPixels", C.c_uint64), ("TilePairs", C.c_uint64),
                ("Tiles", C.c_uint64), ("KernelLaunches", C.c_uint64), ("Reruns", C.c_uint64), ("StoppedObjects", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# edge_info, projekt.h:17-37
EDGE_INFO_DTYPE = np.dtype({
    "names": ["YMax", "XMin", "ZMin", "OneOverZMin", "Gradient", "ZGradient", "OneOverZGradient",
              "YMin", "UMin", "VMin", "UGradient", "VGradient", "Left", "MinColor",
              "ColorGradient", "MinNormal", "NormalGradient", "Next"],
    "formats": ["<i4", "<f4", "<f4", "<f4", "<f4", "<f4", "<f4", "<i4", "<f4", "<f4", "<f4", "<f4",
                "<i4", ("<f4", 4), ("<f4", 4), ("<f4", 3), ("<f4", 3), "<u8"],
    "offsets": [0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 68, 84, 96, 112],
    "itemsize": 120})


class peer_handle(C.Structure):
    """b200r_peer_handle: a CUDA IPC memory handle, 64 opaque bytes any transport can carry."""
    _fields_ = [("Bytes", C.c_ubyte * 64)]


class B200RasterError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"b200r status {code}: {message}")
        self.code = code


_lib = None


def load_library():
    """dlopen libb200raster.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " or `make -C cpu_renderer_b200/csrc` -- there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        ctx = C.c_void_p
        lib.b200r_create.argtypes = [C.POINTER(ctx), C.c_int]
        lib.b200r_destroy.argtypes = [ctx]
        lib.b200r_destroy.restype = None
        lib.b200r_last_error.argtypes = [ctx]
        lib.b200r_last_error.restype = C.c_char_p
        lib.b200r_set_stream.argtypes = [ctx, C.c_void_p]
        lib.b200r_sync.argtypes = [ctx]
        lib.b200r_set_tile.argtypes = [ctx, C.c_int, C.c_int]
        lib.b200r_render_objects.argtypes = [ctx, C.POINTER(render_entry_3d_object), C.c_uint32,
                                             C.POINTER(game_render_commands),
                                             C.POINTER(loaded_bitmap), C.c_uint32]
        lib.b200r_fill_edge_table.argtypes = [ctx, C.POINTER(render_entry_3d_object),
                                              C.POINTER(game_render_commands), C.c_int32]
        lib.b200r_render_device.argtypes = [ctx, C.POINTER(device_mesh), C.c_uint32,
                                            C.POINTER(game_render_commands),
                                            C.POINTER(device_target), C.c_uint32]
        lib.b200r_clear_device.argtypes = [ctx, C.POINTER(device_target), C.c_uint32, C.c_float]
        lib.b200r_get_stats.argtypes = [ctx, C.POINTER(frame_stats)]
        lib.b200r_set_profiling.argtypes = [ctx, C.c_int]
        lib.b200r_get_stage_ms.argtypes = [ctx, C.POINTER(C.c_float)]
        lib.b200r_set_gather_target.argtypes = [ctx, C.POINTER(device_target)]
        lib.b200r_peer_alloc.argtypes = [ctx, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(peer_handle)]
        lib.b200r_peer_open.argtypes = [ctx, C.POINTER(peer_handle), C.POINTER(C.c_void_p)]
        lib.b200r_peer_release.argtypes = [ctx, C.c_void_p]
        _lib = lib
    return _lib


def make_commands(scene, zbuffer_ptr=None, zstride=None):
    """game_render_commands for a cpu_renderer_b200.scene.Scene; returns (struct, keepalive)."""
    n = len(scene.lights)
    lights = (light_info * max(n, 1))()
    for i, l in enumerate(scene.lights):
        lights[i].P = v3(*l.P)
        lights[i].Intensity = v4(*l.intensity)
    cmd = game_render_commands()
    cmd.Width = zstride if zstride is not None else scene.width
    cmd.ZBuffer = zbuffer_ptr
    cmd.LightData.AmbientIntensity = v4(*scene.ambient)
    cmd.LightData.LightCount = n
    cmd.LightData.Lights = C.cast(lights, C.POINTER(light_info))
    t = scene.transform
    cmd.Transform.MetersToPixels = t.meters_to_pixels
    cmd.Transform.ScreenCenter = v2(*t.screen_center)
    cmd.Transform.FocalLength = t.focal_length
    cmd.Transform.DistanceAboveTarget = t.distance_above_target
    return cmd, lights


class Renderer:
    """One context per GPU (b200r_create)."""

    def __init__(self, device: int = -1):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.b200r_create(C.byref(self.ctx), device)
        if rc != OK:
            raise B200RasterError(rc, "b200r_create failed (no B200 / CUDA device?)")

    def close(self):
        if self.ctx:
            self.lib.b200r_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise B200RasterError(rc, self.lib.b200r_last_error(self.ctx).decode())
        return rc

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.b200r_set_stream(self.ctx, C.c_void_p(cuda_stream)))

    def set_tile(self, w: int, h: int):
        self._check(self.lib.b200r_set_tile(self.ctx, w, h))

    def sync(self):
        self._check(self.lib.b200r_sync(self.ctx))

    def set_profiling(self, enable: bool):
        self._check(self.lib.b200r_set_profiling(self.ctx, 1 if enable else 0))

    def stage_ms(self) -> dict:
        ms = (C.c_float * len(STAGES))()
        self._check(self.lib.b200r_get_stage_ms(self.ctx, ms))
        return dict(zip(STAGES, [float(x) for x in ms]))

    def stats(self) -> dict:
        s = frame_stats()
        self._check(self.lib.b200r_get_stats(self.ctx, C.byref(s)))
        return s.as_dict()

    # ---- host-pointer drop-in (FillEdgeTable + DrawModel pair, projekt.cpp:3882 + 162) ----
    def render_scene_host(self, scene, color: np.ndarray, depth: np.ndarray, splits=None, flags=0, phong=False,
                          textured=None, object_ps=None):
        """Render ``scene`` into host arrays color[H,W] u32 / depth[H,W] f32 in place.
        ``splits``: optional list of vertex counts to submit the scene as several objects.
        ``phong``: bool, or one bool per split (render_entry_3d_object::PhongShading).
        ``textured``: bool, or one bool per split: the object carries scene.texture as its Bitmap
        (default: every object, when the scene has a texture).
        ``object_ps``: optional list of Object->P; the whole scene is then submitted once per entry
        (the same vertex arrays as several objects at different positions)."""
        assert color.dtype == np.uint32 and depth.dtype == np.float32
        nv = scene.positions.shape[0]
        splits = splits or [nv]
        assert sum(splits) == nv
        if object_ps is not None:
            assert splits == [nv]
            splits = [nv] * len(object_ps)
        objs = (render_entry_3d_object * len(splits))()
        tex = getattr(scene, "texture", None)
        if textured is None:
            textured = tex is not None
        tex_bmp = None
        if tex is not None:
            tex = np.ascontiguousarray(tex, dtype=np.uint32)
            tex_bmp = loaded_bitmap(tex.shape[1], tex.shape[0], tex.strides[0], tex.ctypes.data)
        at = 0
        for i, cnt in enumerate(splits):
            o = objs[i]
            o.P = v3(*(object_ps[i] if object_ps is not None else scene.object_p))
            if object_ps is not None:
                at = 0
            o.VertexCount = cnt
            o.PhongShading = int(phong[i] if isinstance(phong, (list, tuple)) else phong)
            o.VertexData = scene.positions.ctypes.data + at * 12
            o.ColorData = scene.colors.ctypes.data + at * 16
            o.NormalData = scene.normals.ctypes.data + at * 12
            o.UVData = scene.uvs.ctypes.data + at * 8
            if tex_bmp is not None and (textured[i] if isinstance(textured, (list, tuple)) else textured):
                o.Bitmap = C.addressof(tex_bmp)
            at += cnt
        cmd, keep = make_commands(scene, depth.ctypes.data, depth.strides[0] // 4)
        bmp = loaded_bitmap(color.shape[1], color.shape[0], color.strides[0], color.ctypes.data)
        self._check(self.lib.b200r_render_objects(self.ctx, objs, len(splits), C.byref(cmd),
                                                  C.byref(bmp), flags))
        del keep

    def fill_edge_table(self, scene, first_vertex=0, vertex_count=None, phong=False):
        """b200r_fill_edge_table over (a slice of) the scene as ONE object -> edge_info array."""
        if vertex_count is None:
            vertex_count = scene.positions.shape[0] - first_vertex
        edges = np.zeros(max(vertex_count, 1), dtype=EDGE_INFO_DTYPE)
        o = render_entry_3d_object()
        o.P = v3(*scene.object_p)
        o.VertexCount = vertex_count
        o.VertexData = scene.positions.ctypes.data + first_vertex * 12
        o.ColorData = scene.colors.ctypes.data + first_vertex * 16
        o.NormalData = scene.normals.ctypes.data + first_vertex * 12
        o.UVData = scene.uvs.ctypes.data + first_vertex * 8
        o.EdgeMemory = edges.ctypes.data
        tex = getattr(scene, "texture", None)
        if tex is not None:
            tex = np.ascontiguousarray(tex, dtype=np.uint32)
            tex_bmp = loaded_bitmap(tex.shape[1], tex.shape[0], tex.strides[0], tex.ctypes.data)
            o.Bitmap = C.addressof(tex_bmp)
        cmd, keep = make_commands(scene)
        o.PhongShading = 1 if phong else 0
        n = self._check(self.lib.b200r_fill_edge_table(self.ctx, C.byref(o), C.byref(cmd), 1 if phong else 0))
        del keep
        return edges[:n].copy()

    # ---- device-resident path -------------------------------------------------------------
    def render_device(self, meshes, cmd, target: device_target, flags=0):
        arr = (device_mesh * len(meshes))(*meshes)
        self._check(self.lib.b200r_render_device(self.ctx, arr, len(meshes), C.byref(cmd),
                                                 C.byref(target), flags))

    def clear_device(self, target: device_target, color: int, depth: float):
        self._check(self.lib.b200r_clear_device(self.ctx, C.byref(target), color, depth))

    # ---- fused gather: bands / frames mirrored into one (peer-mapped) image of the whole screen ----
    def set_gather_target(self, target):
        """b200r_set_gather_target; ``None`` switches the mirror off."""
        self._check(self.lib.b200r_set_gather_target(self.ctx, C.byref(target) if target is not None else None))

    def peer_alloc(self, nbytes: int):
        """Device memory other processes can map: returns (device pointer, 64-byte handle)."""
        ptr, h = C.c_void_p(), peer_handle()
        self._check(self.lib.b200r_peer_alloc(self.ctx, nbytes, C.byref(ptr), C.byref(h)))
        return ptr.value, bytes(h.Bytes)

    def peer_open(self, handle: bytes) -> int:
        h = peer_handle()
        C.memmove(h.Bytes, handle, 64)
        ptr = C.c_void_p()
        self._check(self.lib.b200r_peer_open(self.ctx, C.byref(h), C.byref(ptr)))
        return ptr.value

    def peer_release(self, ptr: int):
        self._check(self.lib.b200r_peer_release(self.ctx, C.c_void_p(ptr)))

"""The C-ABI library loads and exports every symbol include/b200_raster.h declares.  No compute
calls here (no GPU in the build container); without a device every entry point must refuse."""
import ctypes as C
import os
import re

import pytest

from cpu_renderer_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_raster.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200r_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_layouts_match_projekt_h():
    # sizes from the reference build: projekt.h:2-15 and 17-37
    assert C.sizeof(api.render_entry_3d_object) == 72
    assert api.EDGE_INFO_DTYPE.itemsize == 120
    assert api.EDGE_INFO_DTYPE.fields["YMin"][1] == 28 and api.EDGE_INFO_DTYPE.fields["Left"][1] == 48
    assert api.EDGE_INFO_DTYPE.fields["MinColor"][1] == 52 and api.EDGE_INFO_DTYPE.fields["Next"][1] == 112
    assert C.sizeof(api.loaded_bitmap) == 24 and C.sizeof(api.game_render_commands) == 104


# ctypes mirror -> C type in include/b200_raster.h
MIRRORS = {
    "render_entry_3d_object": api.render_entry_3d_object, "loaded_bitmap": api.loaded_bitmap,
    "game_render_commands": api.game_render_commands, "light_data": api.light_data, "light_info": api.light_info,
    "projective_transform": api.projective_transform,
    "b200r_device_mesh": api.device_mesh, "b200r_device_texture": api.device_texture,
    "b200r_device_target": api.device_target, "b200r_frame_stats": api.frame_stats,
    "b200r_peer_handle": api.peer_handle,
}


def test_ctypes_mirrors_match_the_header_field_by_field(tmp_path):
    """Compile a C program against the real header and compare sizeof and every field offset of the
    public structs with the hand-written ctypes mirrors in cpu_renderer_b200/api.py."""
    import subprocess
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for cname, mirror in MIRRORS.items():
        lines.append(f'  printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe), str(src)])
    got = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        cname, field, value = line.split()
        got[(cname, field)] = int(value)
    for cname, mirror in MIRRORS.items():
        assert got[(cname, "sizeof")] == C.sizeof(mirror), cname
        for fname, _ in mirror._fields_:
            assert got[(cname, fname)] == getattr(mirror, fname).offset, (cname, fname)
    assert api.EDGE_INFO_DTYPE.itemsize == 120


def test_stage_count_matches_the_header():
    text = open(HEADER).read()
    assert int(re.search(r"#define\s+B200R_STAGES\s+(\d+)", text).group(1)) == len(api.STAGES)
    assert int(re.search(r"#define\s+B200R_WHOLE_OBJECT_AEL\s+(\d+)u", text).group(1)) == api.WHOLE_OBJECT_AEL
    assert int(re.search(r"#define\s+B200R_MESH_PHONG\s+(\d+)u", text).group(1)) == api.MESH_PHONG


def test_no_device_means_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = api.load_library()
    ctx = C.c_void_p()
    assert lib.b200r_create(C.byref(ctx), 0) == api.E_NO_DEVICE
    assert not ctx.value
    with pytest.raises(api.B200RasterError):
        api.Renderer(0)


def test_null_context_is_rejected_not_dereferenced():
    lib = api.load_library()
    assert lib.b200r_sync(None) == api.E_INVALID
    assert lib.b200r_set_tile(None, 64, 32) == api.E_INVALID
    assert lib.b200r_render_objects(None, None, 0, None, None, 0) == api.E_INVALID
    assert lib.b200r_fill_edge_table(None, None, None, 0) == api.E_INVALID
    assert lib.b200r_render_device(None, None, 0, None, None, 0) == api.E_INVALID
    assert lib.b200r_get_stats(None, None) == api.E_INVALID
    assert lib.b200r_set_gather_target(None, None) == api.E_INVALID
    assert lib.b200r_peer_alloc(None, 16, None, None) == api.E_INVALID
    assert lib.b200r_peer_open(None, None, None) == api.E_INVALID
    assert lib.b200r_peer_release(None, None) == api.E_INVALID
    assert lib.b200r_last_error(None) == b"null context"
    lib.b200r_destroy(None)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cpu_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "raster_oracle" not in text, f
                assert "libprojekt_ref" not in text, f

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

"""The C++ host mirror (cpu_renderer_b200/host/b200_dropin.hpp) driven the way the reference's
render-group walker would: a compiled C++ program calls FillEdgeTable + DrawModel through the
C ABI; its image hashes must equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from cpu_renderer_b200 import scene as sc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_program_matches_oracle(tmp_path):
    exe = os.path.join(ROOT, "tests", "host_dropin_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host_dropin_test.cpp"),
                           "-L" + os.path.join(ROOT, "cpu_renderer_b200"), "-lb200raster",
                           "-Wl,-rpath," + os.path.join(ROOT, "cpu_renderer_b200")])
    s = sc.triangle_soup("cpp", 0x123, 20_000, 800, 450, 2.0, 30.0)
    payload = s.positions.tobytes() + s.colors.tobytes() + s.normals.tobytes()
    out = subprocess.run([exe, "0", str(s.triangle_count), str(s.width), str(s.height)], input=payload,
                         capture_output=True, check=True).stdout.decode().split()
    want = ol.oracle_render(s)
    edges, n = ol.oracle_edge_table(s)
    assert int(out[0]) == n
    assert out[1] == ol.fnv1a64_words(want["color"])
    assert out[2] == ol.fnv1a64_words(want["z"])

"""The product's MergeSort key (cpu_renderer_b200/csrc/merge_order.h, compiled into edge_table_kernels.cu) is
pinned on the CPU: sorting by (YMin, b200r_merge_tie_path) must reproduce the permutations of the verbatim
reference's MergeSort (projekt.cpp:2-72; tests/golden/reference_vectors.npz, generated from oracle/_ref) and of
the oracle's restatement on keys with many ties.  The header is compiled for the host with g++; no GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="module")
def tie_path(tmp_path_factory):
    d = tmp_path_factory.mktemp("merge_order")
    src = d / "tie.cpp"
    src.write_text('#include "merge_order.h"\n'
                   'extern "C" void tie_paths(unsigned n, unsigned *out) '
                   '{ for(unsigned i = 0; i < n; ++i) out[i] = b200r_merge_tie_path(i, n); }\n')
    so = d / "libtie.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "cpu_renderer_b200", "csrc"),
                           str(src), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.tie_paths.argtypes = [C.c_uint32, C.c_void_p]

    def f(n):
        out = np.zeros(max(n, 1), dtype=np.uint32)
        lib.tie_paths(n, out.ctypes.data)
        return out[:n]
    return f


def order_by_key(ymin, paths):
    key = ((ymin.astype(np.int64) + 2**31).astype(np.uint64) << np.uint64(32)) | paths.astype(np.uint64)
    assert len(np.unique(key)) == len(key)              # unique: any correct sort gives the same permutation
    return np.argsort(key, kind="stable")


def test_key_order_equals_the_verbatim_merge_sort_permutations(tie_path):
    keys, perm, at = GOLD["mergesort_keys"], GOLD["mergesort_perm"], 0
    for n in GOLD["mergesort_sizes"]:
        n = int(n)
        got = order_by_key(keys[at:at + n], tie_path(n))
        assert np.array_equal(got, perm[at:at + n]), n
        at += n


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 31, 33, 100, 683, 2047, 2048, 2049, 4097, 50_001])
def test_key_order_equals_the_oracle_merge_sort_on_heavy_ties(tie_path, n):
    lib = ol.oracle()
    rng = np.random.default_rng(n)
    e = np.zeros(n, dtype=ol.ORC_EDGE_DTYPE)
    e["YMin"] = rng.integers(-3, 9, size=n)              # a dozen distinct rows: nearly every comparison is a tie
    e["Triangle"] = np.arange(n)
    ymin = e["YMin"].copy()
    tmp = np.zeros(n, dtype=ol.ORC_EDGE_DTYPE)
    lib.orc_merge_sort(n, e.ctypes.data, tmp.ctypes.data)
    assert np.array_equal(order_by_key(ymin, tie_path(n)), e["Triangle"]), n

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/abx.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "abx.h")).read()
    return sorted(set(re.findall(r"^ABX_API [^;(]*?\b(abx_\w+)\(", h, re.M)))


def test_header_declares_entry_points():
    names = _declared()
    assert "abx_bvh_build" in names and "abx_query_spatial_crs" in names and "abx_dbscan" in names
    assert len(names) >= 25


def test_library_exports_every_declared_symbol():
    from arborx_b200 import _lib
    L = _lib.lib()
    for name in _declared():
        assert hasattr(L, name), name
    assert set(_declared()) == set(_lib.SIGNATURES), "Python binding and header disagree"
    assert L.abx_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from arborx_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    st = L.abx_bvh_build(None, 0, None, 0, C.byref(h))
    assert st == _lib.ABX_ERR_CUDA
    assert b"no CPU fallback" in L.abx_last_error()
    import arborx_b200 as abx
    with pytest.raises(RuntimeError):
        abx.ExecutionSpace()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "arborx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert "liborc" not in src and "arborx_oracle" not in src, f

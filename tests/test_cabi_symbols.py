"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/*.h declares; compute calls fail loudly (non-zero status + message) when no CUDA device is present."""
import ctypes as C
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cornerstone-octree_b200", "cstone_b200", "libcstone_b200.so")


def declared_symbols():
    names = set()
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(hdr).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build the library first: python __graft_entry__.py"
    lib = C.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this check is for machines without a GPU")
    lib = C.CDLL(LIB)
    lib.cs_last_error.restype = C.c_char_p
    lim = (C.c_double * 6)(0, 1, 0, 1, 0, 1)
    bnd = (C.c_int * 3)(0, 0, 0)
    buf = (C.c_double * 8)()
    keys = (C.c_uint64 * 8)()
    st = lib.cs_compute_sfc_keys_u64d(0, buf, buf, buf, keys, C.c_size_t(8), lim, bnd, None)
    assert st != 0, "compute call must fail without a CUDA device"
    assert lib.cs_last_error()


def test_python_layer_requires_the_extension(monkeypatch):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "cornerstone-octree_b200"))
    from cstone_b200 import capi

    monkeypatch.setattr(capi, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    monkeypatch.setattr(capi, "_lib", None)
    with pytest.raises(capi.CstoneError):
        capi.lib()

"""CPU: the C-ABI library loads and exports every symbol include/*.h declares;
the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as ct
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from ldpc_sparc_b200 import _lib, ldpc


def _declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", text):
            names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert {"sumprod", "sumprod2", "minsum", "Lxor", "Lxfb", "lsb_ldpc_decode_dev_f32"} <= declared
    handle = ct.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(handle, n)]
    assert not missing, "declared in include/ but not exported: %s" % missing
    # and the ctypes signature table covers exactly the header
    assert set(_lib.SIGNATURES) == declared


def test_version_and_error_slot():
    L = _lib.lib()
    assert L.lsb_version() == 1
    assert isinstance(_lib.last_error(), str)
    assert L.lsb_device_count() >= 0


def test_invalid_arguments_are_reported_not_thrown():
    L = _lib.lib()
    h = ct.c_void_p()
    proto = np.full((2, 4), -1, dtype=np.int32)   # rows of degree 0: invalid
    rc = L.lsb_ldpc_graph_create_qc(proto.ctypes.data_as(_lib.c_i32p), 2, 4, 5, ct.byref(h))
    assert rc == -1 and "degree" in _lib.last_error()
    assert L.lsb_ldpc_graph_create_qc(None, 2, 4, 5, ct.byref(h)) == -1


def test_no_cpu_fallback_without_gpu():
    if _lib.lib().lsb_device_count() > 0:
        pytest.skip("a GPU is visible")
    c = ldpc.code("802.11n", "1/2", 27)
    with pytest.raises(_lib.LibraryError):
        c.decode(np.ones(c.N))
    with pytest.raises(_lib.LibraryError):
        c.Lxor(1.0, 2.0)

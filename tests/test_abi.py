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


def test_row_gather_schedule_is_conflict_free():
    """lsb_sparc_schedule_row_gather (host only): the AMP kernel's row-gather order is a proper edge colouring of
    (residual bank pair, buffer bank pair): every row once, no two rows of a half-warp share a bank pair on either side."""
    import ctypes as ct
    import numpy as np
    from ldpc_sparc_b200 import _lib
    L = _lib.lib()
    rng = np.random.RandomState(7)
    for n_rows, row0 in ((3413, 0), (3413, 3413), (1000, 5), (17, 3), (4095, 1)):
        ids = rng.permutation(4096)[:n_rows].astype(np.uint16)
        words = np.zeros(16384 * 2, dtype=np.uint32)
        chunks = ct.c_int(0)
        _lib.check(L.lsb_sparc_schedule_row_gather(ids.ctypes.data, n_rows, row0, words.ctypes.data, words.size,
                                                   ct.byref(chunks)), "schedule")
        w = words[:chunks.value * 256 * 4].reshape(chunks.value, 16, 16, 4)     # [chunk][half-warp][lane][slot]
        live = w != 0xFFFFFFFF
        rows, slots = (w & 0xFFFF).astype(np.int64), (w >> 16).astype(np.int64)
        assert np.array_equal(np.sort(rows[live]), np.arange(n_rows))            # every row exactly once
        assert np.array_equal(slots[live], ids[rows[live]])                      # with its own slot
        for c in range(chunks.value):
            for hw in range(16):
                for s in range(4):
                    m = live[c, hw, :, s]
                    rb, sb = (row0 + rows[c, hw, m, s]) % 16, slots[c, hw, m, s] % 16
                    assert len(set(rb)) == len(rb) and len(set(sb)) == len(sb)
        degree = max(np.bincount((row0 + np.arange(n_rows)) % 16, minlength=16).max(), np.bincount(ids % 16, minlength=16).max())
        assert chunks.value == (((degree + 15) // 16) + 3) // 4                   # Delta colours: no slot wasted

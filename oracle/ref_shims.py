"""Import the UNMODIFIED reference (/root/reference) under a documented shim set.

ORACLE / TEST INFRASTRUCTURE ONLY.  Usable only in the build container (the
reference tree does not exist on the GPU box).  Used by oracle/make_golden.py
to generate tests/golden/*.npz and by the container-only tests that validate
the numpy/C restatements against the real reference.

Shims, none of which edits a reference file:
  1. ``ctypes.CDLL("C:/msys64/.../c_ldpc.dll")`` (ldpc_jossy/py/ldpc.py:467,493,498)
     is redirected to oracle/_ref/c_ldpc_ref.so, i.e. the reference's own
     c_ldpc.c compiled in place by oracle/Makefile.
  2. ``code.decode`` passes int32 ``intrlv`` as ``long*`` (ldpc.py:395,480); on
     LP64 Linux that reads out of bounds, so graph arrays are widened to C long
     after construction.
  3. numpy>=1.24 removed ``np.float/np.complex/np.object`` (sparc.py:415,463,465,
     803-804,916; sparc_sim.py:197,200): aliases restored.
  4. numpy>=2 turns every non-empty string into True in ``astype('bool')``, which
     breaks ``int_2_bin_arr`` (sparc.py:191-197, sparc_new.py:1372-1378): replaced
     by a function with the numpy-1.x meaning ('0' -> False, '1' -> True).
  5. Import layout: sparc_public modules import each other flat; sparc_sophie
     imports ``ldpc_jossy.py.ldpc`` with the reference root on sys.path.
"""
import ctypes
import os
import sys

import numpy as np

REF_ROOT = os.environ.get("LDPC_SPARC_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "c_ldpc_ref.so")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "sparc_public")) and os.path.exists(REF_SO)


def _int_2_bin_arr(integer, arr_length):
    assert integer >= 0
    return np.array([ch == "1" for ch in np.binary_repr(integer, arr_length)], dtype=bool)


_loaded = {}


def load():
    """Returns a dict of reference modules: sparc, sparc_sim, ldpc, sparc_new, sparc_sim_new."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree or oracle/_ref/c_ldpc_ref.so not present")

    # shim 3
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, typ in (("float", float), ("complex", complex), ("object", object)):
            if name not in vars(np):
                setattr(np, name, typ)

    # shim 1
    real_cdll = ctypes.CDLL

    class _CDLL(real_cdll):
        def __init__(self, name, *a, **k):
            if isinstance(name, str) and name.endswith("c_ldpc.dll"):
                name = REF_SO
            super().__init__(name, *a, **k)

    ctypes.CDLL = _CDLL

    # shim 5
    for p in (os.path.join(REF_ROOT, "sparc_public"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)

    import sparc as ref_sparc  # noqa: E402  (sparc_public/sparc.py)
    import sparc_sim as ref_sparc_sim  # noqa: E402
    from ldpc_jossy.py import ldpc as ref_ldpc  # noqa: E402
    from sparc_sophie import sparc_new as ref_sparc_new  # noqa: E402
    from sparc_sophie import sparc_sim_new as ref_sparc_sim_new  # noqa: E402

    # shim 4
    ref_sparc.int_2_bin_arr = _int_2_bin_arr
    ref_sparc_new.int_2_bin_arr = _int_2_bin_arr

    # shim 2
    orig_init = ref_ldpc.code.__init__

    def _init(self, *a, **k):
        orig_init(self, *a, **k)
        self.vdeg = np.ascontiguousarray(self.vdeg, dtype=np.int_)
        self.cdeg = np.ascontiguousarray(self.cdeg, dtype=np.int_)
        self.intrlv = np.ascontiguousarray(self.intrlv, dtype=np.int_)

    ref_ldpc.code.__init__ = _init

    _loaded.update(
        sparc=ref_sparc,
        sparc_sim=ref_sparc_sim,
        ldpc=ref_ldpc,
        sparc_new=ref_sparc_new,
        sparc_sim_new=ref_sparc_sim_new,
    )
    return _loaded

"""ctypes wrapper of the C oracle (oracle/ldpc_oracle.c) and of the compiled
reference (oracle/_ref/c_ldpc_ref.so).  ORACLE / TEST INFRASTRUCTURE ONLY: may be
imported from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs,
never from ldpc_sparc_b200/.
"""
import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_ref", "libldpc_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "c_ldpc_ref.so")

_f64p = ct.POINTER(ct.c_double)
_i64p = ct.POINTER(ct.c_int64)
_i32p = ct.POINTER(ct.c_int32)
_longp = ct.POINTER(ct.c_long)


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = ct.CDLL(ORACLE_SO)
        sig = [_f64p, _i64p, _i64p, _i64p, ct.c_int64, ct.c_int64, ct.c_int64, _f64p, ct.c_int]
        L.orc_sumprod2.argtypes = sig
        L.orc_sumprod.argtypes = sig
        L.orc_minsum.argtypes = sig[:-1] + [ct.c_double, ct.c_int]
        L.orc_lxor.argtypes = [ct.c_double, ct.c_double, ct.c_int]
        L.orc_lxor.restype = ct.c_double
        L.orc_lxfb.argtypes = [_f64p, ct.c_int64, ct.c_int]
        L.orc_lxfb.restype = ct.c_double
        L.orc_sumprod2_batch.argtypes = [_f64p, _i64p, _i64p, _i64p, ct.c_int64, ct.c_int64, ct.c_int64,
                                         _f64p, _i32p, ct.c_int64, ct.c_int]
        L.orc_sumprod2_batch.restype = None
        _oracle = L
    return _oracle


def ref_lib():
    """The reference's own c_ldpc.c, compiled (None if it was never built)."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        L = ct.CDLL(REF_SO)
        sig = [_f64p, _longp, _longp, _longp, ct.c_int, ct.c_int, ct.c_int, _f64p, ct.c_int]
        L.sumprod2.argtypes = sig
        L.sumprod.argtypes = sig
        L.Lxor.argtypes = [ct.c_double, ct.c_double, ct.c_int]
        L.Lxor.restype = ct.c_double
        L.Lxfb.argtypes = [_f64p, ct.c_long, ct.c_int]
        L.Lxfb.restype = ct.c_double
        _ref = L
    return _ref


class Graph:
    """(vdeg, cdeg, intrlv) as contiguous int64 (== C long on LP64)."""

    def __init__(self, vdeg, cdeg, intrlv):
        self.vdeg = np.ascontiguousarray(vdeg, dtype=np.int64)
        self.cdeg = np.ascontiguousarray(cdeg, dtype=np.int64)
        self.intrlv = np.ascontiguousarray(intrlv, dtype=np.int64)
        self.nv, self.nc, self.nmsg = len(self.vdeg), len(self.cdeg), len(self.intrlv)

    def _args(self, tp):
        return (self.vdeg.ctypes.data_as(tp), self.cdeg.ctypes.data_as(tp), self.intrlv.ctypes.data_as(tp))


def decode(graph, ch, max_it, algo="sumprod2", scale=0.7, impl="oracle"):
    """One frame through the oracle restatement (impl='oracle') or the compiled
    reference (impl='ref').  Returns (app float64[nv], iterations)."""
    ch = np.ascontiguousarray(ch, dtype=np.float64)
    assert ch.shape == (graph.nv,)
    app = np.zeros(graph.nv, dtype=np.float64)
    if impl == "oracle":
        L = oracle_lib()
        a = (ch.ctypes.data_as(_f64p),) + graph._args(_i64p) + (graph.nv, graph.nc, graph.nmsg, app.ctypes.data_as(_f64p))
        if algo == "minsum":
            it = L.orc_minsum(*a, float(scale), int(max_it))
        else:
            it = getattr(L, "orc_" + algo)(*a, int(max_it))
    else:
        L = ref_lib()
        if L is None:
            raise RuntimeError("oracle/_ref/c_ldpc_ref.so not built")
        a = (ch.ctypes.data_as(_f64p),) + graph._args(_longp) + (graph.nv, graph.nc, graph.nmsg, app.ctypes.data_as(_f64p))
        it = getattr(L, algo)(*a, int(max_it))
    return app, it


def decode_batch(graph, ch, max_it, impl="oracle"):
    """sumprod2 over rows of ch[F, nv]; returns (app[F, nv], iters int32[F])."""
    ch = np.ascontiguousarray(ch, dtype=np.float64)
    F = ch.shape[0]
    app = np.zeros_like(ch)
    iters = np.zeros(F, dtype=np.int32)
    if impl == "oracle":
        oracle_lib().orc_sumprod2_batch(ch.ctypes.data_as(_f64p), *graph._args(_i64p), graph.nv, graph.nc,
                                        graph.nmsg, app.ctypes.data_as(_f64p), iters.ctypes.data_as(_i32p), F,
                                        int(max_it))
    else:
        for f in range(F):
            app[f], iters[f] = decode(graph, ch[f], max_it, impl="ref")
    return app, iters


def lxor(a, b, corr=1):
    return oracle_lib().orc_lxor(float(a), float(b), int(corr))


def lxfb(L, corr=1):
    L = np.array(L, dtype=np.float64)
    out = oracle_lib().orc_lxfb(L.ctypes.data_as(_f64p), len(L), int(corr))
    return out, L

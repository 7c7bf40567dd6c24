"""ctypes binding of libldpcsparc_b200.so (see include/ldpc_sparc_b200.h).

There is NO CPU fallback: if the shared library is missing, or no CUDA device is
visible when a compute entry point is called, an exception is raised.
"""
import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LDPC_SPARC_B200_LIB: alternative build of the same library (A/B kernel experiments)
LIB_PATH = os.environ.get("LDPC_SPARC_B200_LIB") or os.path.join(_HERE, "libldpcsparc_b200.so")

c_f32p = ct.POINTER(ct.c_float)
c_f64p = ct.POINTER(ct.c_double)
c_i32p = ct.POINTER(ct.c_int32)
c_i64p = ct.POINTER(ct.c_int64)
c_u8p = ct.POINTER(ct.c_uint8)
c_u32p = ct.POINTER(ct.c_uint32)
c_longp = ct.POINTER(ct.c_long)


class LibraryError(RuntimeError):
    pass


class AmpParams(ct.Structure):      # lsb_sparc_amp_params
    _fields_ = [("awgn_var", ct.c_double), ("rtol", ct.c_double), ("atol", ct.c_double),
                ("t_max", ct.c_int32), ("phi_est_method", ct.c_int32), ("modulation", ct.c_int32),
                ("reserved", ct.c_int32)]


class AmpOut(ct.Structure):         # lsb_sparc_amp_out
    _fields_ = [("idx", ct.c_void_p), ("t_final", ct.c_void_p), ("psi", ct.c_void_p), ("nmse", ct.c_void_p),
                ("beta_soft", ct.c_void_p), ("s_last", ct.c_void_p), ("tau_trace", ct.c_void_p),
                ("psi_trace", ct.c_void_p)]


_lib = None

# name -> (restype, argtypes); every symbol include/ldpc_sparc_b200.h declares
SIGNATURES = {
    "lsb_version": (ct.c_int, []),
    "lsb_last_error": (ct.c_char_p, []),
    "lsb_device_count": (ct.c_int, []),
    "lsb_launch_count": (ct.c_int64, []),
    "lsb_launch_count_reset": (None, []),
    "lsb_host_alloc": (ct.c_void_p, [ct.c_uint64]),
    "lsb_host_free": (None, [ct.c_void_p]),
    # legacy (c_ldpc.c)
    "sumprod": (ct.c_int, [c_f64p, c_longp, c_longp, c_longp, ct.c_int, ct.c_int, ct.c_int, c_f64p, ct.c_int]),
    "sumprod2": (ct.c_int, [c_f64p, c_longp, c_longp, c_longp, ct.c_int, ct.c_int, ct.c_int, c_f64p, ct.c_int]),
    "minsum": (ct.c_int, [c_f64p, c_longp, c_longp, c_longp, ct.c_int, ct.c_int, ct.c_int, c_f64p, ct.c_double, ct.c_int]),
    "Lxor": (ct.c_double, [ct.c_double, ct.c_double, ct.c_int]),
    "Lxfb": (ct.c_double, [c_f64p, ct.c_long, ct.c_int]),
    # batched LDPC
    "lsb_ldpc_graph_create_qc": (ct.c_int, [c_i32p, ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_void_p)]),
    "lsb_ldpc_graph_create": (ct.c_int, [c_i64p, c_i64p, c_i64p, ct.c_int64, ct.c_int64, ct.c_int64, ct.POINTER(ct.c_void_p)]),
    "lsb_ldpc_graph_destroy": (None, [ct.c_void_p]),
    "lsb_ldpc_graph_info": (ct.c_int, [ct.c_void_p, c_i64p]),
    "lsb_ldpc_mc_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_double, ct.c_uint64, ct.c_uint64, ct.c_int64,
                                       ct.c_int, ct.c_int, ct.c_double, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p,
                                       ct.c_void_p]),
    "lsb_ldpc_encode_dev": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_ldpc_decode_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, ct.c_double, ct.c_void_p]),
    "lsb_ldpc_decode_dev_f64": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, ct.c_double, ct.c_void_p]),
    "lsb_ldpc_decode_host_bits_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, ct.c_double]),
    "lsb_ldpc_decode_host_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, ct.c_double]),
    "lsb_ldpc_decode_host_f64": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, ct.c_double, ct.c_int]),
    # SPARC (sub-sampled DCT design)
    "lsb_sparc_design_create": (ct.c_int, [c_f64p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, c_u32p, c_u32p, ct.POINTER(ct.c_void_p)]),
    "lsb_sparc_design_destroy": (None, [ct.c_void_p]),
    "lsb_band_design_create": (ct.c_int, [c_f64p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, c_u32p, c_u32p, ct.POINTER(ct.c_void_p)]),
    "lsb_band_design_destroy": (None, [ct.c_void_p]),
    "lsb_band_design_bytes": (ct.c_int64, [ct.c_void_p]),
    "lsb_band_ab_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_band_az_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_band_amp_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.POINTER(AmpParams), ct.c_void_p, ct.POINTER(AmpOut), ct.c_void_p]),
    "lsb_sparc_schedule_row_gather": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, ct.c_void_p, ct.c_int64, ct.POINTER(ct.c_int)]),
    "lsb_sparc_design_info": (ct.c_int, [ct.c_void_p, c_i64p]),
    "lsb_sparc_ab_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_sparc_ab_dev_f64": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_sparc_az_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_sparc_az_dev_f64": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_sparc_amp_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.POINTER(AmpParams), ct.c_void_p, ct.POINTER(AmpOut), ct.c_void_p]),
    "lsb_sparc_amp_dev_f64": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.POINTER(AmpParams), ct.c_void_p, ct.POINTER(AmpOut), ct.c_void_p]),
    "lsb_sparc_amp_host_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.POINTER(AmpParams), ct.c_void_p, ct.c_void_p, ct.c_void_p]),
    "lsb_band_amp_host_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.POINTER(AmpParams), ct.c_void_p, ct.c_void_p, ct.c_void_p]),
    # SPARC (dense Gaussian design) + section <-> bit glue
    "lsb_dense_design_create": (ct.c_int, [c_f64p, ct.c_int64, ct.c_int64, ct.c_int, ct.c_int, ct.POINTER(ct.c_void_p)]),
    "lsb_dense_design_destroy": (None, [ct.c_void_p]),
    "lsb_tc_design_create": (ct.c_int, [c_f64p, ct.c_int64, ct.c_int64, ct.POINTER(ct.c_void_p)]),
    "lsb_tc_design_destroy": (None, [ct.c_void_p]),
    "lsb_tc_ab_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_tc_az_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_tc_amp_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, ct.c_double, ct.c_double,
                                      ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p]),
    "lsb_tc_amp_host_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, ct.c_double, ct.c_double,
                                       ct.c_int, ct.c_void_p]),
    # complex SPARCs (sub-sampled DFT design)
    "lsb_csparc_design_create": (ct.c_int, [c_f64p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, c_u32p, c_u32p, ct.POINTER(ct.c_void_p)]),
    "lsb_csparc_design_destroy": (None, [ct.c_void_p]),
    "lsb_csparc_ab_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_csparc_ab_dev_f64": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_csparc_az_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_csparc_az_dev_f64": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p]),
    "lsb_csparc_amp_dev_f32": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.POINTER(AmpParams), ct.c_void_p, ct.POINTER(AmpOut), ct.c_void_p]),
    "lsb_csparc_amp_dev_f64": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int64, ct.POINTER(AmpParams), ct.c_void_p, ct.POINTER(AmpOut), ct.c_void_p]),
    # state evolution
    "lsb_sparc_se_expect_f64": (ct.c_int, [ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, c_f64p, ct.c_int, c_f64p, ct.c_void_p]),
    "lsb_sparc_se_expect_f32": (ct.c_int, [ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, c_f64p, ct.c_int, c_f64p, ct.c_void_p]),
    "lsb_sparc_se_expect_c128": (ct.c_int, [ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, c_f64p, ct.c_int, c_f64p, ct.c_void_p]),
    "lsb_sparc_se_expect_c64": (ct.c_int, [ct.c_void_p, ct.c_int64, ct.c_int, ct.c_int, c_f64p, ct.c_int, c_f64p, ct.c_void_p]),
}
_VP = ct.c_void_p
for _suf in ("f32", "f64"):
    SIGNATURES["lsb_dense_ab_dev_" + _suf] = (ct.c_int, [_VP, _VP, _VP, ct.c_int64, _VP])
    SIGNATURES["lsb_dense_az_dev_" + _suf] = (ct.c_int, [_VP, _VP, _VP, ct.c_int64, _VP])
    SIGNATURES["lsb_amp_residual_dev_" + _suf] = (ct.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, ct.c_int64, ct.c_int, ct.c_double, ct.c_int, ct.c_int, _VP])
    SIGNATURES["lsb_amp_denoise_dev_" + _suf] = (ct.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, ct.c_int64, ct.c_int, ct.c_int, ct.c_double, _VP])
    SIGNATURES["lsb_section_to_bit_dev_" + _suf] = (ct.c_int, [_VP, _VP, _VP, ct.c_int64, ct.c_int, ct.c_double, _VP])
    SIGNATURES["lsb_bit_to_section_dev_" + _suf] = (ct.c_int, [_VP, _VP, _VP, _VP, ct.c_int64, ct.c_int, ct.c_int, ct.c_double, _VP])
    SIGNATURES["lsb_hard_info_bits_dev_" + _suf] = (ct.c_int, [_VP, _VP, ct.c_int64, ct.c_int, ct.c_int, _VP])
    SIGNATURES["lsb_integrated_onsager_dev_" + _suf] = (ct.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, ct.c_int64, ct.c_int, ct.c_int, ct.c_double, _VP])


def lib():
    """Loads the library once; raises LibraryError if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or ./build.sh (there is no CPU fallback)" % LIB_PATH
            )
        handle = ct.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error():
    return lib().lsb_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise LibraryError("%s failed (status %d): %s" % (what, rc, last_error()))


def require_device():
    if lib().lsb_device_count() < 1:
        raise LibraryError("no CUDA device visible: the decoders run on the GPU only (no CPU fallback)")


ALGO = {"sumprod2": 0, "sumprod": 1, "minsum": 2}

"""SPARC golden vectors from the UNMODIFIED reference (see make_golden.py).

sparc_public: for every case the reference's sparc_encode -> awgn_channel ->
sparc_decode chain is run under oracle/ref_shims.py, with
msg_vector_mmse_estimator wrapped (not edited) to record the per-iteration
(s, tau, beta).  Stored per case/seed: code/decode params, awgn_var, seed, bits_in,
x, y, and the reference's bits_out, section decisions, t_final, nmse, final psi,
per-iteration tau / psi, and the last iteration's s and soft beta.
"""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SPARC_CASES = {
    # BASELINE config 1: power-allocated, L=1024 M=32 R=1.5 (n=3413, w=4096, 32 blocks)
    "c1_pa": (dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9), dict(t_max=25), 1.0),
    # regular (single block, w=8192)
    "reg": (dict(P=15.0, R=1.3, L=256, M=32), dict(t_max=25), 1.0),
    # spatially coupled (9x8 band matrix, 16 non-zero blocks, w=2048)
    "sc": (dict(P=15.0, R=1.4, L=256, M=32, spatially_coupled=True, omega=2, Lambda=8), dict(t_max=25), 1.0),
    # SC + PA, M=16 (< warp), residual-variance estimate from z (phi_est_method=2), low SNR (errors)
    "sc_pa_m16": (dict(P=15.0, R=1.4, L=256, M=16, spatially_coupled=True, omega=3, Lambda=8, power_allocated=True,
                       B=2, R_PA_ratio=0.5), dict(t_max=20, phi_est_method=2), 1.0),
    # M=64 (> warp), tighter rtol
    "reg_m64": (dict(P=7.0, R=1.0, L=128, M=64), dict(t_max=30, rtol=1e-4), 1.0),
    # transform size 65536 (does not fit shared memory): regular, one block of 384 x 32768
    "big_reg": (dict(P=15.0, R=1.5, L=64, M=512), dict(t_max=20), 1.0),
    # BASELINE config 4: spatially coupled "wave" decoding, sparc_demo_sc_decode_wave.ipynb:60-68
    # (Lr=37, Lc=32, n=12284, w=65536, 192 blocks); one seed, no soft-state arrays (4 MB each)
    "c4_sc_wave": (dict(P=15.0, R=1.5, L=2048, M=512, spatially_coupled=True, omega=6, Lambda=32), dict(t_max=40), 1.0),
}
BIG = {"c4_sc_wave"}
SEEDS = ([1, 2], [5, 6])


def run_reference(ref, cp, dp, var, seed):
    R, RS = ref["sparc"], ref["sparc_sim"]
    cp, dp = dict(cp), dict(dp)
    bits_i, beta0, x, Ab, Az = R.sparc_encode(cp, var, seed)
    y = RS.awgn_channel(x, var, seed)
    trace = []
    orig = R.msg_vector_mmse_estimator

    def spy(s, tau, M, K=1):
        out = orig(s, tau, M, K)
        trace.append((s.copy(), np.copy(tau), out.copy()))
        return out

    R.msg_vector_mmse_estimator = spy
    try:
        bits_o, beta, T, nmse, expect = R.sparc_decode(y, cp, dp, var, seed, beta0, Ab, Az)
    finally:
        R.msg_vector_mmse_estimator = orig
    L, M = cp["L"], cp["M"]
    W = R.create_base_matrix(**dict(cp, awgn_var=var))
    Lc = 1 if W.ndim == 0 else W.shape[-1]
    Mc = L * M // Lc
    tau_it = np.array([np.broadcast_to(np.asarray(t[1], dtype=float).reshape(-1), (L * M,))[::Mc] if np.ndim(t[1]) else
                       np.full(Lc, float(t[1])) for t in trace])
    psi_it = np.array([1 - (t[2] ** 2).reshape(Lc, -1).sum(axis=1) / (L / Lc) for t in trace])
    return dict(bits_in=bits_i, x=x, y=y, bits_out=bits_o, idx=beta.reshape(L, M).argmax(1).astype(np.int32),
                t_final=np.array(T), nmse=np.asarray(nmse), expect_err=np.array(bool(expect)),
                tau_it=tau_it, psi_it=psi_it, s_last=trace[-1][0].astype(np.float32),
                beta_soft=trace[-1][2].astype(np.float32), n=np.array(cp["n"]))


def make_sparc(ref):
    out = {}
    meta = {}
    for name, (cp, dp, var) in SPARC_CASES.items():
        meta[name] = dict(code_params=cp, decode_params=dp, awgn_var=var, seeds=[list(s) for s in SEEDS])
        seeds = SEEDS[:1] if name in BIG else SEEDS
        meta[name]["seeds"] = [list(s) for s in seeds]
        for si, seed in enumerate(seeds):
            r = run_reference(ref, cp, dp, var, seed)
            logM = int(np.log2(cp["M"]))
            truth = r["bits_in"].reshape(-1, logM).astype(int).dot(1 << np.arange(logM)[::-1])
            print(name, seed, "t_final", int(r["t_final"]), "section errors", int((r["idx"] != truth).sum()), "n", int(r["n"]))
            if name in BIG:
                for k in ("s_last", "beta_soft", "x"):
                    r.pop(k)
                r["bits_in"] = np.packbits(r["bits_in"])
                r["bits_out"] = np.packbits(r["bits_out"])
            for k, v in r.items():
                out["%s_s%d_%s" % (name, si, k)] = v
    out["meta_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "sparc_reference.npz"), **out)
    print("sparc_reference.npz: %.1f kB" % (os.path.getsize(os.path.join(GOLD, "sparc_reference.npz")) / 1e3))


def make_sophie(ref):
    try:
        from make_golden_sophie import make_sophie as _mk
    except ImportError:
        print("sophie goldens: generator not present yet")
        return
    _mk(ref)


# ------------------------------------------------------------------ real K = 2 modulation
# Separate file (tests/golden/sparc_mod_reference.npz) so the unmodulated goldens above stay
# byte-identical:  python oracle/make_golden_sparc.py mod
MOD_CASES = {
    "k2_reg": (dict(P=15.0, R=1.3, L=256, M=32, modulated=True, K=2), dict(t_max=25), 1.0),
    "k2_pa_m64": (dict(P=15.0, R=1.5, L=256, M=64, modulated=True, K=2, power_allocated=True, B=8, R_PA_ratio=0.9),
                  dict(t_max=25), 1.0),
    # M < warp, SC, residual-variance estimate from z, rate close to capacity (section errors)
    "k2_sc_m16": (dict(P=15.0, R=1.55, L=256, M=16, modulated=True, K=2, spatially_coupled=True, omega=3, Lambda=8),
                  dict(t_max=20, phi_est_method=2), 1.0),
}


def make_sparc_mod(ref):
    out, meta = {}, {}
    for name, (cp, dp, var) in MOD_CASES.items():
        meta[name] = dict(code_params=cp, decode_params=dp, awgn_var=var, seeds=[list(s) for s in SEEDS])
        for si, seed in enumerate(SEEDS):
            r = run_reference(ref, cp, dp, var, seed)
            L, M = cp["L"], cp["M"]
            sec = int(np.log2(M)) + 1
            b = r["bits_in"].reshape(L, sec)
            truth = b[:, :-1].astype(int).dot(1 << np.arange(sec - 1)[::-1])
            dense = ref["sparc"].bin_arr_2_msg_vector(r["bits_out"], M, 2).reshape(L, M)
            r["idx"] = np.abs(dense).argmax(1).astype(np.int32)                 # decided location ...
            r["val"] = dense[np.arange(L), r["idx"]].astype(np.int8)            # ... and value (+1 / -1)
            errs = int(((r["idx"] != truth) | ((r["val"] < 0) != b[:, -1])).sum())
            print(name, seed, "t_final", int(r["t_final"]), "section errors", errs, "n", int(r["n"]))
            for k, v in r.items():
                out["%s_s%d_%s" % (name, si, k)] = v
    out["meta_json"] = np.array(json.dumps(meta))
    path = os.path.join(GOLD, "sparc_mod_reference.npz")
    np.savez_compressed(path, **out)
    print("sparc_mod_reference.npz: %.1f kB" % (os.path.getsize(path) / 1e3))


# ------------------------------------------------------------------ complex designs
# tests/golden/sparc_complex_reference.npz:  python oracle/make_golden_sparc.py complex
COMPLEX_CASES = {
    "c_k1": (dict(P=15.0, R=2.6, L=256, M=32, complex=True), dict(t_max=25), 1.0),
    "c_k2": (dict(P=15.0, R=2.4, L=128, M=64, complex=True, modulated=True, K=2), dict(t_max=25), 1.0),
    "c_k4_pa": (dict(P=15.0, R=3.0, L=256, M=32, complex=True, modulated=True, K=4, power_allocated=True, B=8,
                     R_PA_ratio=0.9), dict(t_max=25), 1.0),
    # K > 4 branch, M < warp, SC, residual-variance estimate from z, rate near capacity (section errors)
    "c_k8_sc_m16": (dict(P=15.0, R=3.3, L=256, M=16, complex=True, modulated=True, K=8, spatially_coupled=True,
                         omega=3, Lambda=8), dict(t_max=20, phi_est_method=2), 1.0),
}


def symbols_of(ref, beta, L, M, K):
    """Decided / true section symbols as location + M * constellation index."""
    dense = beta.reshape(L, M)
    loc = np.abs(dense).argmax(1)
    if K == 1:
        return loc.astype(np.int32)
    c = ref["sparc"].psk_constel(K)
    k = np.array([int(np.flatnonzero(c == v)[0]) for v in dense[np.arange(L), loc]])
    return (loc + M * k).astype(np.int32)


def run_reference_complex(ref, cp, dp, var, seed):
    R, RS = ref["sparc"], ref["sparc_sim"]
    cp, dp = dict(cp), dict(dp)
    bits_i, beta0, x, Ab, Az = R.sparc_encode(cp, var, seed)
    y = RS.awgn_channel(x, var, seed)
    trace = []
    orig = R.msg_vector_mmse_estimator

    def spy(s, tau, M, K=1):
        tau_in = np.copy(tau)                  # the estimator halves a complex design's tau in place (:417-418)
        out = orig(s, tau, M, K)
        trace.append((s.copy(), tau_in, out.copy()))
        return out

    R.msg_vector_mmse_estimator = spy
    try:
        bits_o, beta, T, nmse, expect = R.sparc_decode(y, cp, dp, var, seed, beta0, Ab, Az)
    finally:
        R.msg_vector_mmse_estimator = orig
    L, M = cp["L"], cp["M"]
    K = cp["K"] if cp["modulated"] else 1
    W = R.create_base_matrix(**dict(cp, awgn_var=var))
    Lc = 1 if W.ndim == 0 else W.shape[-1]
    Mc = L * M // Lc
    tau_it = np.array([np.broadcast_to(np.asarray(t[1], dtype=float).reshape(-1), (L * M,))[::Mc] if np.ndim(t[1]) else
                       np.full(Lc, float(t[1])) for t in trace])
    psi_it = np.array([1 - (np.abs(t[2]) ** 2).reshape(Lc, -1).sum(axis=1) / (L / Lc) for t in trace])
    return dict(bits_in=bits_i, x=x, y=y, bits_out=bits_o, sym=symbols_of(ref, beta, L, M, K),
                sym0=symbols_of(ref, beta0, L, M, K), t_final=np.array(T), nmse=np.asarray(nmse),
                expect_err=np.array(bool(expect)), tau_it=tau_it, psi_it=psi_it,
                s_last=trace[-1][0].astype(np.complex64), beta_soft=trace[-1][2].astype(np.complex64),
                n=np.array(cp["n"]))


def make_sparc_complex(ref):
    out, meta = {}, {}
    for name, (cp, dp, var) in COMPLEX_CASES.items():
        meta[name] = dict(code_params=cp, decode_params=dp, awgn_var=var, seeds=[list(s) for s in SEEDS])
        for si, seed in enumerate(SEEDS):
            r = run_reference_complex(ref, cp, dp, var, seed)
            print(name, seed, "t_final", int(r["t_final"]), "section errors", int((r["sym"] != r["sym0"]).sum()),
                  "bit errors", int((r["bits_in"] != r["bits_out"]).sum()), "n", int(r["n"]))
            for k, v in r.items():
                out["%s_s%d_%s" % (name, si, k)] = v
    out["meta_json"] = np.array(json.dumps(meta))
    path = os.path.join(GOLD, "sparc_complex_reference.npz")
    np.savez_compressed(path, **out)
    print("sparc_complex_reference.npz: %.1f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import ref_shims
    if sys.argv[1:] == ["mod"]:
        make_sparc_mod(ref_shims.load())
    elif sys.argv[1:] == ["complex"]:
        make_sparc_complex(ref_shims.load())
    else:
        make_sparc(ref_shims.load())

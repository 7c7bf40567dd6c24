"""GPU parity tests of the SPARC path (sub-sampled DCT design): Ab/Az operators and
the AMP decoder against the reference's golden vectors and the numpy oracle.

Tolerances: fp64 mode -- per-iteration tau / psi to 1e-9 relative, t_final and
section decisions exact; fp32 mode (throughput) -- tau / psi / beta within 1e-4
relative (north_star), section decisions exact.
"""
import json
import os

import numpy as np
import pytest

import sparc_oracle as so
from conftest import GOLDEN
from ldpc_sparc_b200 import sparc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "sparc_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


# ------------------------------------------------------------------ operators
@pytest.mark.parametrize("shape", ["scalar", "vector", "matrix"])
def test_ab_az_against_dense_matrix(shape):
    rng = np.random.RandomState(1)
    if shape == "scalar":
        W, L, M, n = np.array(3.0), 16, 8, 70
    elif shape == "vector":
        W, L, M, n = np.array([1.0, 2.5, 0.5, 4.0]), 16, 16, 100
    else:
        W, L, M, n = so.sc_basic(np.array([2.0, 1.0]), 2, 3), 12, 8, 48
    Ab, Az = sparc.sparc_transforms(W, L, M, n, [3, 4])
    A = so.Design(W, L, M, n, [3, 4]).dense()
    x, y = rng.randn(5, L * M), rng.randn(5, n)
    for prec, tol in (("fp64", 1e-11), ("fp32", 2e-5)):
        ax = Ab(x, precision=prec)
        ay = Az(y, precision=prec)
        assert ax.shape == (5, n) and ay.shape == (5, L * M)
        assert np.max(np.abs(ax - x @ A.T)) <= tol * np.max(np.abs(x @ A.T))
        assert np.max(np.abs(ay - y @ A)) <= tol * np.max(np.abs(y @ A))
    # single vectors, as the reference closures are called
    assert np.allclose(Ab(x[0]), A @ x[0], rtol=1e-10, atol=1e-10)
    assert np.allclose(Az(y[0]), A.T @ y[0], rtol=1e-10, atol=1e-10)
    # adjointness <Ax, y> == <x, A^T y>
    assert np.isclose(np.dot(Ab(x[1]), y[1]), np.dot(x[1], Az(y[1])), rtol=1e-10)


def test_operators_c1_size_vs_oracle(gold):
    """BASELINE config 1 design (w = 4096, 32 blocks): encode x = A beta0 equals the
    reference's codeword, A^T y equals the oracle's."""
    g, meta = gold
    m = meta["c1_pa"]
    cp = dict(m["code_params"])
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], m["seeds"][0])
    assert cp["n"] == 3413 and np.array_equal(bits_in, g["c1_pa_s0_bits_in"])
    assert rel_err(x, g["c1_pa_s0_x"]) < 1e-9 or np.max(np.abs(x - g["c1_pa_s0_x"])) < 1e-10
    y = g["c1_pa_s0_y"]
    d = so.Design(sparc.create_base_matrix(**dict(cp, awgn_var=1.0)), 1024, 32, 3413, m["seeds"][0])
    ref = d.Az(y)
    assert np.max(np.abs(Az(y) - ref)) < 1e-10 * np.max(np.abs(ref))
    assert np.max(np.abs(Az(y, precision="fp32") - ref)) < 2e-5 * np.max(np.abs(ref))


# ---------------------------------------------------------------- AMP goldens
def _decode_case(g, meta, name, si, precision, **kw):
    m = meta[name]
    pre = "%s_s%d_" % (name, si)
    cp, dp = dict(m["code_params"]), dict(m["decode_params"])
    seed = m["seeds"][si]
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
    sparc.check_decode_params(dp)
    y = g[pre + "y"]
    out = sparc.sparc_amp(y, cp, dp, m["awgn_var"], seed, beta0, Ab, Az, precision=precision,
                          return_soft=True, return_trace=True)
    return pre, cp, dp, out


@pytest.mark.parametrize("name", ["c1_pa", "reg", "sc", "sc_pa_m16", "reg_m64", "big_reg"])
def test_amp_fp64_matches_reference_golden(gold, name):
    g, meta = gold
    for si in range(2):
        pre, cp, dp, (beta, t_final, nmse, psi, ex) = _decode_case(g, meta, name, si, "fp64")
        T = int(g[pre + "t_final"])
        assert t_final == T
        assert np.array_equal(beta.reshape(-1, cp["M"]).argmax(1), g[pre + "idx"])
        assert rel_err(ex["tau"][:T], g[pre + "tau_it"]) < 1e-9
        ref_psi = g[pre + "psi_it"]
        assert np.max(np.abs(ex["psi"][:T] - ref_psi)) <= 1e-9 * np.max(np.abs(ref_psi)) + 1e-13
        assert np.max(np.abs(ex["s"] - g[pre + "s_last"])) <= 2e-6 * np.max(np.abs(g[pre + "s_last"]))   # golden s is fp32
        assert np.allclose(nmse, g[pre + "nmse"], rtol=1e-5, atol=1e-7)                                      # nmse out is fp32


@pytest.mark.parametrize("name", ["c1_pa", "reg", "sc", "sc_pa_m16", "reg_m64", "big_reg"])
def test_amp_fp32_matches_reference_golden(gold, name):
    g, meta = gold
    for si in range(2):
        pre, cp, dp, (beta, t_final, nmse, psi, ex) = _decode_case(g, meta, name, si, "fp32")
        T = int(g[pre + "t_final"])
        assert np.array_equal(beta.reshape(-1, cp["M"]).argmax(1), g[pre + "idx"])       # decisions exact
        # The reference stops when every psi_c moved by < 1e-6 relative (sparc.py:984);
        # fp32 iterates reach that only once they are bitwise stationary, which can
        # take a few more iterations (never fewer): decisions are unaffected.
        assert T - 1 <= t_final <= dp["t_max"] - 1, (t_final, T)
        Tc = min(T, t_final)
        assert rel_err(ex["tau"][:Tc], g[pre + "tau_it"][:Tc]) < 1e-4
        ref_psi = g[pre + "psi_it"][:Tc]
        # psi within 1e-4 relative, plus an absolute floor: once a block has converged
        # (psi ~ 1e-6) psi = sum beta(1-beta) is exp(-gap/tau) with gap/tau ~ 10..100, so
        # the fp32 rounding of s is amplified by 1/tau; tau (the quantity the decoder
        # uses) stays within 1e-4 because it depends on sum_c W_c psi_c only
        assert np.all(np.abs(ex["psi"][:Tc] - ref_psi) <= 1e-4 * ref_psi + 2e-8)
        if t_final == T:
            ref_beta = g[pre + "beta_soft"].astype(np.float64)
            assert np.max(np.abs(ex["beta_soft"] - ref_beta)) < 1e-4


def test_sparc_decode_dropin_single_frame(gold):
    """sparc_encode -> awgn -> sparc_decode exactly as sparc_sim.py:21-24 calls them."""
    g, meta = gold
    m = meta["c1_pa"]
    for si in range(2):
        pre = "c1_pa_s%d_" % si
        cp, dp = dict(m["code_params"]), dict(m["decode_params"])
        seed = m["seeds"][si]
        bits_i, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
        y = so.awgn_channel(x, m["awgn_var"], seed)
        assert np.max(np.abs(y - g[pre + "y"])) < 1e-9
        bits_o, beta, T, nmse, expect = sparc.sparc_decode(g[pre + "y"], cp, dp, m["awgn_var"], seed, beta0, Ab, Az,
                                                           precision="fp64")
        assert bits_o.dtype == bool and np.array_equal(bits_o, g[pre + "bits_out"])
        assert beta.shape == (1024 * 32,) and beta.sum() == 1024 and T == int(g[pre + "t_final"])
        assert nmse.shape == (25, 32) and bool(expect) == bool(g[pre + "expect_err"])
        # decoder that rebuilds the transforms from rand_seed (Ab = Az = None, sparc.py:910-911)
        bits_o2, _, T2, _, _ = sparc.sparc_decode(g[pre + "y"], cp, dp, m["awgn_var"], seed, beta0, precision="fp64")
        assert np.array_equal(bits_o2, bits_o) and T2 == T


def test_batched_decode_vs_oracle():
    """64 frames sharing one design (C1 shape), messages and noise per frame: section
    decisions identical to the oracle's frame by frame, in both precisions."""
    import torch
    cp = dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9)
    dp = dict(t_max=25)
    sparc.check_code_params(cp)
    sparc.check_decode_params(dp)
    W = sparc.create_base_matrix(**dict(cp, awgn_var=1.0))
    n = 3413
    cp["n"] = n
    Ab, Az = sparc.sparc_transforms(W, 1024, 32, n, [1, 2])
    rng = np.random.RandomState(0)
    F = 64
    idx0 = rng.randint(0, 32, (F, 1024))
    beta0 = np.zeros((F, 1024 * 32))
    beta0[np.arange(F)[:, None], np.arange(1024)[None, :] * 32 + idx0] = 1
    y = Ab(beta0) + rng.randn(F, n)
    d = so.Design(W, 1024, 32, n, [1, 2])
    ref = [so.sparc_decode(y[f], cp, dp, 1.0, d, beta0[f]) for f in range(8)]
    for prec in ("fp64", "fp32"):
        bits, idx, T, nmse, expect = sparc.sparc_decode(y, cp, dp, 1.0, [1, 2], idx0, Ab, Az, precision=prec)
        assert idx.shape == (F, 1024) and bits.shape == (F, 5120) and nmse.shape == (F, 25, 32)
        for f in range(8):
            assert np.array_equal(idx[f], ref[f][1].reshape(1024, 32).argmax(1)), (prec, f)
            assert np.array_equal(bits[f], ref[f][0])
            if prec == "fp64":
                assert T[f] == ref[f][2]
                assert np.allclose(nmse[f], ref[f][3], rtol=1e-5, atol=1e-7)
        assert (idx != idx0).mean() < 5e-3                       # section error rate of this operating point
    # device-tensor path returns tensors and the same decisions
    yt = torch.from_numpy(y).cuda()
    bits_t, idx_t, T_t, _, _ = sparc.sparc_decode(yt, cp, dp, 1.0, [1, 2], None, Ab, Az)
    assert np.array_equal(idx_t.cpu().numpy(), idx)
    assert np.array_equal(bits_t.cpu().numpy(), bits)


def _c1_batch(F, seed=0):
    cp = dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9)
    dp = dict(t_max=25)
    sparc.check_code_params(cp)
    sparc.check_decode_params(dp)
    W = sparc.create_base_matrix(**dict(cp, awgn_var=1.0))
    cp["n"] = 3413
    Ab, Az = sparc.sparc_transforms(W, 1024, 32, 3413, [1, 2])
    rng = np.random.RandomState(seed)
    idx0 = rng.randint(0, 32, (F, 1024))
    beta0 = np.zeros((F, 1024 * 32), dtype=np.float32)
    beta0[np.arange(F)[:, None], np.arange(1024)[None, :] * 32 + idx0] = 1
    y = (Ab(beta0, precision="fp32") + rng.randn(F, 3413)).astype(np.float32)
    return cp, dp, Ab, Az, idx0, y


def test_host_entry_equals_device_path_and_is_reproducible():
    """lsb_sparc_amp_host_f32 (host buffers in / out) returns exactly what the device-pointer call returns;
    two runs of the same batch are bit-identical (t_final, decisions, psi): the psi / nmse sums are taken in
    a fixed order, no atomics (scheduling-dependent early stop)."""
    import torch
    cp, dp, Ab, Az, idx0, y = _c1_batch(37)                  # odd frame count: the last CTA holds one frame only
    yt = torch.from_numpy(y).cuda()
    outs = []
    for _ in range(2):
        idx_d, T_d, _, psi_d = sparc.sparc_amp(yt, cp, dp, 1.0, [1, 2], None, Ab, Az, precision="fp32")
        outs.append((idx_d.cpu().numpy().copy(), T_d.cpu().numpy().copy(), psi_d.cpu().numpy().copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2]), "psi differs between two runs of the same batch"
    idx_h, T_h, psi_h = sparc.sparc_decode_host(y, cp, dp, 1.0, Ab)
    assert np.array_equal(idx_h, outs[0][0]) and np.array_equal(T_h, outs[0][1]) and np.array_equal(psi_h, outs[0][2])
    assert (idx_h != idx0).mean() < 5e-3
    # a frame decoded alone, or next to a different neighbour, gives the same result as inside the batch
    perm = np.arange(37)[::-1].copy()
    idx_p, T_p, psi_p = sparc.sparc_decode_host(y[perm], cp, dp, 1.0, Ab)
    assert np.array_equal(idx_p, idx_h[perm]) and np.array_equal(T_p, T_h[perm]) and np.array_equal(psi_p, psi_h[perm])
    idx_1, T_1, psi_1 = sparc.sparc_decode_host(y[5:6], cp, dp, 1.0, Ab)
    assert np.array_equal(idx_1[0], idx_h[5]) and T_1[0] == T_h[5] and np.array_equal(psi_1[0], psi_h[5])


def test_pair_kernel_against_one_frame_kernel(tmp_path):
    """The two-frames-per-CTA fp32 kernel against the one-frame fp32 kernel (LSB_AMP_PAIR=0; the switch is read
    once per process, so the one-frame numbers come from a subprocess fed the same received words)."""
    import subprocess
    import sys
    cp, dp, Ab, Az, idx0, y = _c1_batch(16, seed=3)
    idx_h, T_h, psi_h = sparc.sparc_decode_host(y, cp, dp, 1.0, Ab)
    np.save(tmp_path / "y.npy", y)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np, sys; sys.path.insert(0, %r)\n"
        "from ldpc_sparc_b200 import sparc\n"
        "cp = dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9); dp = dict(t_max=25)\n"
        "sparc.check_code_params(cp); sparc.check_decode_params(dp)\n"
        "W = sparc.create_base_matrix(**dict(cp, awgn_var=1.0)); cp['n'] = 3413\n"
        "Ab, Az = sparc.sparc_transforms(W, 1024, 32, 3413, [1, 2])\n"
        "idx, T, psi = sparc.sparc_decode_host(np.load(%r), cp, dp, 1.0, Ab)\n"
        "np.savez(%r, idx=idx, T=T, psi=psi)\n"
    ) % (root, str(tmp_path / "y.npy"), str(tmp_path / "one.npz"))
    env = dict(os.environ, LSB_AMP_PAIR="0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    one = np.load(tmp_path / "one.npz")
    assert np.array_equal(one["idx"], idx_h), "decisions of the two fp32 kernels differ"
    # fp32 iterates differ in the last bits, so t_final may differ by the stop rule's noise; psi agrees to fp32 accuracy
    # (a frame whose psi hovers at the fp32 noise floor meets the 1e-6 relative stop test at a different iteration in
    # the two kernels: same decisions and psi, a few iterations apart)
    dT = np.abs(one["T"].astype(int) - T_h.astype(int))
    assert np.mean(dT == 0) >= 0.75 and dT.max() <= 5
    assert np.allclose(one["psi"], psi_h, rtol=0, atol=1e-6)


def test_alternating_designs_keep_their_shared_memory_opt_in():
    """Two designs of different size share the kernels' template instantiations; the dynamic shared-memory
    opt-in is per (kernel, device) and must never shrink (second design lowered it, the first then
    failed to launch)."""
    big = dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9)
    small = dict(P=15.0, R=1.0, L=64, M=32)
    res = {}
    for rep in range(2):
        for name, cp in (("big", big), ("small", small)):
            cp = dict(cp)
            bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, 1.0, [7, 8])
            y = x + np.random.RandomState(1).randn(cp["n"])
            for prec in ("fp64", "fp32"):
                out = sparc.sparc_decode(y, cp, dict(t_max=10), 1.0, [7, 8], beta0, Ab, Az, precision=prec)
                key = (name, prec)
                if key in res:
                    assert np.array_equal(res[key], out[0])
                res[key] = out[0]


def test_config4_sc_wave_matches_reference_golden(gold):
    """BASELINE config 4 (sparc_demo_sc_decode_wave.ipynb: L=2048, M=512, R=1.5, omega=6,
    Lambda=32, t_max=40 -> Lr=37, Lc=32, n=12284, transform size 65536, 192 blocks).  The FFT
    buffer does not fit shared memory, so this runs the global-buffer variant of the kernels."""
    g, meta = gold
    m = meta["c4_sc_wave"]
    pre = "c4_sc_wave_s0_"
    cp, dp = dict(m["code_params"]), dict(m["decode_params"])
    seed = m["seeds"][0]
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
    assert cp["n"] == 12284 and Ab.design.w == 65536
    assert np.array_equal(np.packbits(bits_in), g[pre + "bits_in"])
    y = so.awgn_channel(x, m["awgn_var"], seed)
    assert np.max(np.abs(y - g[pre + "y"])) < 1e-9 * np.max(np.abs(y))      # encode x = A beta0 on the GPU (fp64)
    sparc.check_decode_params(dp)
    T = int(g[pre + "t_final"])
    for prec in ("fp64", "fp32"):
        beta, t_final, nmse, psi, ex = sparc.sparc_amp(g[pre + "y"], cp, dp, m["awgn_var"], seed, beta0, Ab, Az,
                                                       precision=prec, return_trace=True)
        assert np.array_equal(beta.reshape(-1, 512).argmax(1), g[pre + "idx"]), prec
        if prec == "fp64":
            assert t_final == T
            assert rel_err(ex["tau"][:T], g[pre + "tau_it"]) < 1e-9
            assert np.allclose(nmse, g[pre + "nmse"], rtol=1e-5, atol=1e-7)
        else:
            assert T - 1 <= t_final <= dp["t_max"] - 1
            Tc = min(T, t_final)
            assert rel_err(ex["tau"][:Tc], g[pre + "tau_it"][:Tc]) < 1e-4


def test_config4_fp32_against_fp64_on_a_batch(gold):
    """Config 4, four fresh frames (messages and noise) on the reference design: the fp32 throughput kernel against
    the fp64 kernel, which the test above pins to the reference's golden run -- identical section decisions on every
    frame, tau trace within 1e-4 over the iterations both ran, t_final within the fp32 stop rule's slack."""
    g, meta = gold
    m = meta["c4_sc_wave"]
    cp, dp = dict(m["code_params"]), dict(m["decode_params"])
    seed = m["seeds"][0]
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
    sparc.check_decode_params(dp)
    L, M, n = cp["L"], cp["M"], cp["n"]
    rng = np.random.RandomState(42)
    F = 4
    idx0 = rng.randint(0, M, (F, L))
    b0 = np.zeros((F, L * M))
    b0[np.arange(F)[:, None], np.arange(L)[None, :] * M + idx0] = 1
    y = Ab(b0) + np.sqrt(m["awgn_var"]) * rng.randn(F, n)
    out = {}
    for prec in ("fp64", "fp32"):
        out[prec] = sparc.sparc_amp(y, cp, dp, m["awgn_var"], seed, idx0, Ab, Az, precision=prec, return_trace=True)
    idx64, T64, _, _, ex64 = out["fp64"]
    idx32, T32, _, _, ex32 = out["fp32"]
    idx64, idx32, T64, T32 = (np.asarray(v.cpu() if hasattr(v, "cpu") else v) for v in (idx64, idx32, T64, T32))
    assert np.array_equal(idx64, idx32)
    assert (idx64 != idx0).mean() < 5e-3
    tau64, tau32 = (np.asarray(e["tau"].cpu() if hasattr(e["tau"], "cpu") else e["tau"]) for e in (ex64, ex32))
    for f in range(F):
        assert T64[f] - 1 <= T32[f] <= dp["t_max"] - 1
        Tc = int(min(T64[f], T32[f]))
        assert rel_err(tau32[f, :Tc], tau64[f, :Tc]) < 1e-4


# ---------------------------------------------------------------- real 2-PSK modulation
@pytest.fixture(scope="module")
def gold_mod():
    g = np.load(os.path.join(GOLDEN, "sparc_mod_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


def _decided(beta, M):
    dense = beta.reshape(-1, M)
    idx = np.abs(dense).argmax(1)
    return idx, dense[np.arange(len(idx)), idx]


@pytest.mark.parametrize("name", ["k2_reg", "k2_pa_m64", "k2_sc_m16"])
def test_amp_k2_fp64_matches_reference_golden(gold_mod, name):
    g, meta = gold_mod
    for si in range(2):
        pre, cp, dp, (beta, t_final, nmse, psi, ex) = _decode_case(g, meta, name, si, "fp64")
        T = int(g[pre + "t_final"])
        assert t_final == T
        idx, val = _decided(beta, cp["M"])
        assert np.array_equal(idx, g[pre + "idx"]) and np.array_equal(val, g[pre + "val"])      # locations and signs exact
        assert rel_err(ex["tau"][:T], g[pre + "tau_it"]) < 1e-9
        ref_psi = g[pre + "psi_it"]
        assert np.max(np.abs(ex["psi"][:T] - ref_psi)) <= 1e-9 * np.max(np.abs(ref_psi)) + 1e-13
        assert np.max(np.abs(ex["s"] - g[pre + "s_last"])) <= 2e-6 * np.max(np.abs(g[pre + "s_last"]))
        assert np.max(np.abs(ex["beta_soft"] - g[pre + "beta_soft"])) < 1e-6                        # golden beta is fp32
        assert np.allclose(nmse, g[pre + "nmse"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["k2_reg", "k2_pa_m64", "k2_sc_m16"])
def test_amp_k2_fp32_matches_reference_golden(gold_mod, name):
    g, meta = gold_mod
    for si in range(2):
        pre, cp, dp, (beta, t_final, nmse, psi, ex) = _decode_case(g, meta, name, si, "fp32")
        T = int(g[pre + "t_final"])
        idx, val = _decided(beta, cp["M"])
        assert np.array_equal(idx, g[pre + "idx"]) and np.array_equal(val, g[pre + "val"])
        assert T - 1 <= t_final <= dp["t_max"] - 1, (t_final, T)
        Tc = min(T, t_final)
        assert rel_err(ex["tau"][:Tc], g[pre + "tau_it"][:Tc]) < 1e-4
        ref_psi = g[pre + "psi_it"][:Tc]
        assert np.all(np.abs(ex["psi"][:Tc] - ref_psi) <= 1e-4 * ref_psi + 2e-8)


def test_k2_dropin_bits_and_sim(gold_mod):
    """sparc_encode -> awgn -> sparc_decode and sparc_sim for a modulated code, single frame
    and batched (bits: logM location bits then the value bit, sparc.py:330-400)."""
    from ldpc_sparc_b200 import sparc_sim
    g, meta = gold_mod
    m = meta["k2_sc_m16"]
    seed = m["seeds"][0]
    pre = "k2_sc_m16_s0_"
    cp, dp = dict(m["code_params"]), dict(m["decode_params"])
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
    assert np.array_equal(bits_in, g[pre + "bits_in"]) and cp["n"] == int(g[pre + "n"])
    assert np.max(np.abs(x - g[pre + "x"])) < 1e-9
    bits_o, beta, T, nmse, expect = sparc.sparc_decode(g[pre + "y"], cp, dp, m["awgn_var"], seed, beta0, Ab, Az,
                                                       precision="fp64")
    assert np.array_equal(bits_o, g[pre + "bits_out"]) and T == int(g[pre + "t_final"])
    assert bool(expect) == bool(g[pre + "expect_err"])
    # batched: three copies of the frame + the other seed's noise realisation
    y2 = np.stack([g[pre + "y"], g[pre + "y"], g["k2_sc_m16_s0_x"] + (g[pre + "y"] - g[pre + "x"])[::-1]])
    bits_b, idx_b, T_b, _, _ = sparc.sparc_decode(y2, cp, dp, m["awgn_var"], seed, beta0, Ab, Az, precision="fp64")
    assert bits_b.shape == (3, bits_in.size) and np.array_equal(bits_b[0], bits_o) and np.array_equal(bits_b[1], bits_o)
    assert np.array_equal(idx_b[0] % cp["M"], g[pre + "idx"]) and np.array_equal(idx_b[0] >= cp["M"], g[pre + "val"] < 0)
    res = sparc_sim.sparc_sim(dict(m["code_params"]), dict(m["decode_params"]), m["awgn_var"], seed, precision="fp64")
    L, M = cp["L"], cp["M"]
    b = g[pre + "bits_in"].reshape(L, -1)
    truth = b[:, :-1].astype(int).dot(1 << np.arange(b.shape[1] - 1)[::-1])
    loc_err = truth != g[pre + "idx"]
    val_err = b[:, -1] != (g[pre + "val"] < 0)
    assert res["num_of_loc_errs"] == int(loc_err.sum()) and res["num_of_sec_errs"] == int((loc_err | val_err).sum())
    assert res["ber"] == np.count_nonzero(g[pre + "bits_in"] ^ g[pre + "bits_out"]) / bits_in.size
    out = sparc_sim.sparc_sim_batch(dict(P=15.0, R=1.3, L=256, M=32, modulated=True, K=2), dict(t_max=25), 1.0, 384,
                                    chunk=128)
    assert out["frames"] == 384 and out["n"] == 1182 and out["ser"] < 2e-3
    assert out["section_errors"] <= out["bit_errors"] <= 6 * out["section_errors"]


# ---------------------------------------------------------------- complex designs (sub-sampled DFT, K-PSK)
@pytest.fixture(scope="module")
def gold_cplx():
    g = np.load(os.path.join(GOLDEN, "sparc_complex_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


def test_complex_operators_match_dense_and_golden(gold_cplx):
    import sparc_oracle as so
    rng = np.random.RandomState(0)
    for W, L, M, n in ((np.array(3.0), 8, 8, 20), (so.sc_basic(np.array(2.0), 2, 3), 6, 8, 32),
                       (np.array([1.0, 2.0, 0.5, 3.0]), 64, 16, 300)):
        d = sparc.Design(W, L, M, n, [3, 4], csparc=True)
        ref = so.Design(W, L, M, n, [3, 4], csparc=True)
        assert d.w == ref.w and np.array_equal(d.order0, ref.order0) and np.array_equal(d.order1, ref.order1)
        x = rng.randn(3, L * M) + 1j * rng.randn(3, L * M)
        y = rng.randn(3, n) + 1j * rng.randn(3, n)
        for prec, tol in (("fp64", 1e-12), ("fp32", 2e-6)):
            ab, az = d.apply(x, False, prec), d.apply(y, True, prec)
            for f in range(3):
                assert np.max(np.abs(ab[f] - ref.Ab(x[f]))) <= tol * np.max(np.abs(ref.Ab(x[f])))
                assert np.max(np.abs(az[f] - ref.Az(y[f]))) <= tol * np.max(np.abs(ref.Az(y[f])))
    g, meta = gold_cplx                                   # encoder output of the reference itself
    m = meta["c_k4_pa"]
    cp = dict(m["code_params"])
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], m["seeds"][0])
    assert np.array_equal(bits_in, g["c_k4_pa_s0_bits_in"]) and cp["n"] == int(g["c_k4_pa_s0_n"])
    assert np.max(np.abs(x - g["c_k4_pa_s0_x"])) < 1e-10 * np.max(np.abs(x))


@pytest.mark.parametrize("name", ["c_k1", "c_k2", "c_k4_pa", "c_k8_sc_m16"])
def test_complex_amp_fp64_matches_reference_golden(gold_cplx, name):
    g, meta = gold_cplx
    for si in range(2):
        pre, cp, dp, (beta, t_final, nmse, psi, ex) = _decode_case(g, meta, name, si, "fp64")
        K = cp["K"] if cp["modulated"] else 1
        T = int(g[pre + "t_final"])
        assert t_final == T
        assert np.array_equal(sparc.msg_vector_2_symbols(beta, cp["M"], K), g[pre + "sym"])     # locations and PSK symbols exact
        assert beta.dtype == (complex if K > 2 else float)
        assert rel_err(ex["tau"][:T], g[pre + "tau_it"]) < 1e-9
        ref_psi = g[pre + "psi_it"]
        assert np.max(np.abs(ex["psi"][:T] - ref_psi)) <= 1e-9 * np.max(np.abs(ref_psi)) + 1e-13
        assert np.max(np.abs(ex["s"] - g[pre + "s_last"])) <= 2e-6 * np.max(np.abs(g[pre + "s_last"]))
        assert np.max(np.abs(ex["beta_soft"] - g[pre + "beta_soft"])) < 1e-6
        assert np.allclose(nmse, g[pre + "nmse"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["c_k1", "c_k2", "c_k4_pa", "c_k8_sc_m16"])
def test_complex_amp_fp32_matches_reference_golden(gold_cplx, name):
    g, meta = gold_cplx
    for si in range(2):
        pre, cp, dp, (beta, t_final, nmse, psi, ex) = _decode_case(g, meta, name, si, "fp32")
        K = cp["K"] if cp["modulated"] else 1
        T = int(g[pre + "t_final"])
        sym = sparc.msg_vector_2_symbols(beta, cp["M"], K)
        if name in ("c_k1", "c_k2"):                      # converged frames: decisions exact
            assert np.array_equal(sym, g[pre + "sym"])
        else:                                             # frames that end above capacity: undecided sections may flip
            assert np.mean(sym != g[pre + "sym"]) < 0.02
        assert T - 1 <= t_final <= dp["t_max"] - 1, (t_final, T)
        Tc = min(T, t_final)
        assert rel_err(ex["tau"][:Tc], g[pre + "tau_it"][:Tc]) < 1e-4
        ref_psi = g[pre + "psi_it"][:Tc]
        assert np.all(np.abs(ex["psi"][:Tc] - ref_psi) <= 1e-4 * ref_psi + 1e-6)


def test_complex_dropin_decode_and_sims(gold_cplx):
    from ldpc_sparc_b200 import sparc_sim
    g, meta = gold_cplx
    for name in ("c_k4_pa", "c_k8_sc_m16"):
        m = meta[name]
        seed = m["seeds"][1]
        pre = "%s_s1_" % name
        cp, dp = dict(m["code_params"]), dict(m["decode_params"])
        bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
        bits_o, beta, T, nmse, expect = sparc.sparc_decode(g[pre + "y"], cp, dp, m["awgn_var"], seed, beta0, Ab, Az,
                                                           precision="fp64")
        assert np.array_equal(bits_o, g[pre + "bits_out"]) and T == int(g[pre + "t_final"])
        assert bool(expect) == bool(g[pre + "expect_err"])
        y2 = np.stack([g[pre + "y"], g[pre + "y"][::-1].copy()])
        bits_b, sym_b, T_b, _, _ = sparc.sparc_decode(y2, cp, dp, m["awgn_var"], seed, beta0, Ab, Az, precision="fp64")
        assert np.array_equal(bits_b[0], bits_o) and np.array_equal(sym_b[0], g[pre + "sym"]) and T_b[0] == T
        res = sparc_sim.sparc_sim(dict(m["code_params"]), dict(m["decode_params"]), m["awgn_var"], seed, precision="fp64")
        assert res["ber"] == np.count_nonzero(g[pre + "bits_in"] ^ g[pre + "bits_out"]) / bits_in.size
        assert res["num_of_sec_errs"] == int((g[pre + "sym"] != g[pre + "sym0"]).sum())
        assert res["num_of_loc_errs"] == int((g[pre + "sym"] % cp["M"] != g[pre + "sym0"] % cp["M"]).sum())
    out = sparc_sim.sparc_sim_batch(dict(P=15.0, R=2.0, L=256, M=32, complex=True, modulated=True, K=4), dict(t_max=25),
                                    1.0, 256, chunk=128)
    assert out["frames"] == 256 and out["ser"] < 5e-3
    assert out["section_errors"] <= out["bit_errors"] <= 7 * out["section_errors"]

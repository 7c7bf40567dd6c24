"""GPU parity tests of the tensor-core form of the sparc_public decoder (csrc/sparc_band_tc.cu): the sub-sampled DCT
design stored explicitly as a block-banded matrix of fp16 hi / lo pairs, A beta and A^T z as tcgen05 GEMMs over the
frame batch (`precision='tc'`).  Checked against the dense matrix of the numpy oracle, the reference's golden runs and
the fp64 FFT kernel.  Bar of this mode: section decisions and frame / section error counts exact, t_final within the
fp32 stop rule's slack, per-iteration tau within 2e-4 relative (kTauTol; measured <= 8e-5 on every golden incl. full-size
config 4) and the converged tau within 1e-4.  Two things made that possible: operands as fp16
pairs (22 mantissa bits instead of the 16 of fp16 pairs) and K accumulated in TMEM only 8 K blocks at a time, the chunks
summed in registers with round-to-nearest -- the tensor core rounds its fp32 accumulator toward zero at every MMA and
that bias, linear in K, was the dominant error (7e-4 in tau at config 4 before)."""
import json
import os

import numpy as np
import pytest

import sparc_oracle as so
from conftest import GOLDEN
from ldpc_sparc_b200 import sparc

pytestmark = pytest.mark.gpu
kTauTol = 2e-4


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "sparc_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("shape", ["scalar", "vector", "matrix", "matrix_ragged"])
def test_band_operators_against_dense_matrix(shape):
    """Ab / Az through the explicit design: every block shape (one block, a row of blocks, a band), row blocks that do
    not fill a 128-row tile, column blocks that do not fill a 64-element K block, frame counts off the 128 tile."""
    rng = np.random.RandomState(1)
    if shape == "scalar":
        W, L, M, n = np.array(3.0), 16, 8, 70
    elif shape == "vector":
        W, L, M, n = np.array([1.0, 2.5, 0.5, 4.0]), 16, 16, 100
    elif shape == "matrix":
        W, L, M, n = so.sc_basic(np.array([2.0, 1.0]), 2, 3), 12, 8, 48
    else:
        W, L, M, n = so.sc_basic(np.array([2.0, 1.0, 0.5]), 3, 5), 45, 32, 1015      # Lr = 7, Lc = 15, Mr = 145, Mc = 96
    Ab, Az = sparc.sparc_transforms(W, L, M, n, [3, 4])
    A = so.Design(W, L, M, n, [3, 4]).dense()
    for F in (5, 131):
        x, y = rng.randn(F, L * M), rng.randn(F, n)
        ax, ay = Ab(x, precision="tc"), Az(y, precision="tc")
        assert ax.shape == (F, n) and ay.shape == (F, L * M)
        assert np.max(np.abs(ax - x @ A.T)) <= 3e-5 * np.max(np.abs(x @ A.T))
        assert np.max(np.abs(ay - y @ A)) <= 3e-5 * np.max(np.abs(y @ A))


def _decode_case(g, meta, name, si, prec, **kw):
    m = meta[name]
    pre = "%s_s%d_" % (name, si)
    cp, dp = dict(m["code_params"]), dict(m["decode_params"])
    seed = m["seeds"][si]
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
    sparc.check_decode_params(dp)
    return pre, cp, dp, sparc.sparc_amp(g[pre + "y"], cp, dp, m["awgn_var"], seed, beta0, Ab, Az, precision=prec,
                                        return_soft=True, return_trace=True, **kw)


@pytest.mark.parametrize("name", ["c1_pa", "reg", "sc", "sc_pa_m16", "reg_m64", "big_reg"])
def test_band_amp_matches_reference_golden(gold, name):
    """The reference's own decodes (goldens made by oracle/make_golden_sparc.py from the unmodified reference): scalar,
    power-allocated and spatially coupled base matrices, M = 16 .. 64.  Same bar as the fp32 FFT kernel."""
    g, meta = gold
    for si in range(2):
        pre, cp, dp, (beta, t_final, nmse, psi, ex) = _decode_case(g, meta, name, si, "tc")
        T = int(g[pre + "t_final"])
        assert np.array_equal(beta.reshape(-1, cp["M"]).argmax(1), g[pre + "idx"])       # decisions exact
        assert T - 1 <= t_final <= dp["t_max"] - 1, (t_final, T)
        Tc = min(T, t_final)
        assert rel_err(ex["tau"][:Tc], g[pre + "tau_it"][:Tc]) < kTauTol
        assert rel_err(ex["tau"][Tc - 1], g[pre + "tau_it"][Tc - 1]) < 1e-4          # converged
        ref_psi = g[pre + "psi_it"][:Tc]
        # psi_c = sum beta (1 - beta) of a nearly converged block is exp(-gap / tau) with gap / tau ~ 10 .. 100, which
        # amplifies the rounding of s by 1 / tau (same floor as the fp32 FFT kernel's test, scaled)
        assert np.all(np.abs(ex["psi"][:Tc] - ref_psi) <= 2e-3 * ref_psi + 1e-6)
        assert np.allclose(nmse[:Tc + 1], g[pre + "nmse"][:Tc + 1], rtol=5e-3, atol=1e-6)
        if t_final == T:
            ref_beta = g[pre + "beta_soft"].astype(np.float64)
            assert np.max(np.abs(ex["beta_soft"] - ref_beta)) < 2e-4


def test_band_amp_batch_equals_fft_kernels():
    """A batch of 150 frames (not a multiple of the 128-frame tile) of a spatially coupled, power-allocated code:
    decisions, t_final slack and tau trace against the fp64 FFT kernel (= the reference's arithmetic)."""
    cp = dict(P=15.0, R=1.3, L=256, M=32, spatially_coupled=True, omega=3, Lambda=8)
    dp = dict(t_max=30)
    awgn = 1.0
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, awgn, [5, 6])
    sparc.check_decode_params(dp)
    L, M, n = cp["L"], cp["M"], cp["n"]
    rng = np.random.RandomState(3)
    F = 150
    idx0 = rng.randint(0, M, (F, L))
    b0 = np.zeros((F, L * M))
    b0[np.arange(F)[:, None], np.arange(L)[None, :] * M + idx0] = 1
    y = Ab(b0) + np.sqrt(awgn) * rng.randn(F, n)
    out = {}
    for prec in ("fp64", "tc"):
        out[prec] = sparc.sparc_amp(y, cp, dp, awgn, [5, 6], idx0, Ab, Az, precision=prec, return_trace=True)
    idx64, T64, nm64, psi64, ex64 = out["fp64"]
    idxtc, Ttc, nmtc, psitc, extc = out["tc"]
    assert np.array_equal(idx64, idxtc)
    assert (idx64 != idx0).mean() < 2e-2
    errs = []
    for f in range(F):
        assert T64[f] - 1 <= Ttc[f] <= dp["t_max"] - 1
        Tc = int(min(T64[f], Ttc[f]))
        errs.append(rel_err(extc["tau"][f, :Tc], ex64["tau"][f, :Tc]))
    # 150 frames x 30 iterations x 8 column blocks: the slowest-converging frames amplify most in their transient
    assert np.percentile(errs, 90) < kTauTol and max(errs) < 2e-3, (np.percentile(errs, 90), max(errs))


def test_band_config4_sc_wave_matches_reference_golden(gold):
    """BASELINE config 4 at full size (L = 2048, M = 512, Lr = 37, Lc = 32, n = 12284, 192 blocks of 332 x 32768 out of
    w = 65536): the reference's golden frame plus three fresh ones through the explicit design (19.6 GB of fp16
    operands): decisions identical to the reference / the fp64 FFT kernel, tau within kTauTol."""
    g, meta = gold
    m = meta["c4_sc_wave"]
    pre = "c4_sc_wave_s0_"
    cp, dp = dict(m["code_params"]), dict(m["decode_params"])
    seed = m["seeds"][0]
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
    sparc.check_decode_params(dp)
    L, M, n = cp["L"], cp["M"], cp["n"]
    T = int(g[pre + "t_final"])
    beta, t_final, nmse, psi, ex = sparc.sparc_amp(g[pre + "y"], cp, dp, m["awgn_var"], seed, beta0, Ab, Az,
                                                   precision="tc", return_trace=True)
    assert np.array_equal(beta.reshape(-1, M).argmax(1), g[pre + "idx"])
    assert T - 1 <= t_final <= dp["t_max"] - 1
    Tc = min(T, t_final)
    assert rel_err(ex["tau"][:Tc], g[pre + "tau_it"][:Tc]) < kTauTol
    rng = np.random.RandomState(42)
    F = 3
    idx0 = rng.randint(0, M, (F, L))
    b0 = np.zeros((F, L * M))
    b0[np.arange(F)[:, None], np.arange(L)[None, :] * M + idx0] = 1
    y = Ab(b0) + np.sqrt(m["awgn_var"]) * rng.randn(F, n)
    idx64, T64, _, _, ex64 = sparc.sparc_amp(y, cp, dp, m["awgn_var"], seed, idx0, Ab, Az, precision="fp64", return_trace=True)
    idxtc, Ttc, _, _, extc = sparc.sparc_amp(y, cp, dp, m["awgn_var"], seed, idx0, Ab, Az, precision="tc", return_trace=True)
    assert np.array_equal(idx64, idxtc)
    for f in range(F):
        assert T64[f] - 1 <= Ttc[f] <= dp["t_max"] - 1
        Tc = int(min(T64[f], Ttc[f]))
        assert rel_err(extc["tau"][f, :Tc], ex64["tau"][f, :Tc]) < kTauTol


def test_band_host_entry_equals_device_path():
    """lsb_band_amp_host_f32 (host y in; decisions, t_final, psi out) against the device-pointer entry."""
    import torch
    cp = dict(P=15.0, R=1.4, L=256, M=32, spatially_coupled=True, omega=2, Lambda=8)
    dp = dict(t_max=25)
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, 1.0, [1, 2])
    sparc.check_decode_params(dp)
    L, M, n = cp["L"], cp["M"], cp["n"]
    rng = np.random.RandomState(9)
    F = 37
    idx0 = rng.randint(0, M, (F, L))
    b0 = np.zeros((F, L * M))
    b0[np.arange(F)[:, None], np.arange(L)[None, :] * M + idx0] = 1
    y = (Ab(b0) + rng.randn(F, n)).astype(np.float32)
    idx_h, T_h, psi_h = sparc.sparc_decode_host(y, cp, dp, 1.0, Ab, precision="tc")
    idx_d, T_d, _, psi_d = sparc.sparc_amp(torch.from_numpy(y).cuda(), cp, dp, 1.0, [1, 2], None, Ab, Az, precision="tc")
    assert np.array_equal(idx_h, idx_d.cpu().numpy()) and np.array_equal(T_h, T_d.cpu().numpy())
    assert np.array_equal(psi_h, psi_d.cpu().numpy())
    assert (idx_h != idx0).mean() < 5e-2


def test_band_config4_batch_decisions_equal_fft_path():
    """Config 4, 48 fresh frames at the bench's operating point: the tensor-core path and the fp32 FFT path make the
    same section decisions on every frame (and so count the same section / frame errors), t_final within one."""
    import torch
    cp = dict(P=15.0, R=1.5, L=2048, M=512, spatially_coupled=True, omega=6, Lambda=32)
    dp = dict(t_max=40)
    bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, 1.0, [1, 2])
    sparc.check_decode_params(dp)
    L, M, n = cp["L"], cp["M"], cp["n"]
    F = 48
    gen = torch.Generator(device="cuda").manual_seed(77)
    idx0 = torch.randint(0, M, (F, L), device="cuda", generator=gen, dtype=torch.int32)
    b0 = torch.zeros((F, L * M), device="cuda")
    b0.view(F, L, M).scatter_(2, idx0.long().unsqueeze(-1), 1.0)
    y = Ab(b0, precision="fp32") + torch.randn((F, n), device="cuda", generator=gen)
    del b0
    idx_tc, T_tc, _, _ = sparc.sparc_amp(y, cp, dp, 1.0, [1, 2], None, Ab, Az, precision="tc")
    idx_32, T_32, _, _ = sparc.sparc_amp(y, cp, dp, 1.0, [1, 2], None, Ab, Az, precision="fp32")
    assert torch.equal(idx_tc, idx_32)
    assert float((idx_tc != idx0).float().mean()) < 1e-3
    assert int((T_tc - T_32).abs().max()) <= 1

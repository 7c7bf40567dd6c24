"""CPU: the numpy SPARC oracle against the reference's golden vectors, the
imported reference (build container only) and the closed-form design matrix."""
import json
import os

import numpy as np
import pytest

import sparc_oracle as so
from conftest import GOLDEN
from ldpc_sparc_b200 import sparc


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "sparc_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


def _run_oracle(cp, dp, var, seed):
    bits_i, beta0, x, design, n = so.sparc_encode(cp, var, seed)
    y = so.awgn_channel(x, var, seed)
    trace = []
    out = so.sparc_decode(y, dict(cp, n=n), dp, var, design, beta0, trace)
    return bits_i, beta0, x, y, design, out, trace


def test_oracle_matches_reference_golden_config4(gold):
    """BASELINE config 4 (SC wave, w = 65536, 192 blocks, n = 12284): ~20 s of numpy."""
    g, meta = gold
    m = meta["c4_sc_wave"]
    pre = "c4_sc_wave_s0_"
    bits_i, beta0, x, y, design, (bits_o, beta, T, nmse, expect), trace = _run_oracle(
        m["code_params"], m["decode_params"], m["awgn_var"], m["seeds"][0])
    assert (design.Lr, design.Lc, design.w, design.n) == (37, 32, 65536, 12284)
    assert np.array_equal(np.packbits(bits_i), g[pre + "bits_in"]) and np.array_equal(y, g[pre + "y"])
    assert T == int(g[pre + "t_final"]) and np.array_equal(np.packbits(bits_o), g[pre + "bits_out"])
    assert np.array_equal(beta.reshape(-1, 512).argmax(1), g[pre + "idx"])
    assert np.array_equal(nmse, g[pre + "nmse"])
    assert np.array_equal(np.array([t["tau"] for t in trace]), g[pre + "tau_it"])
    assert np.array_equal(np.array([t["psi"] for t in trace]), g[pre + "psi_it"])


def test_oracle_matches_reference_golden(gold):
    g, meta = gold
    for name, m in meta.items():
        if name == "c4_sc_wave":
            continue
        for si, seed in enumerate(m["seeds"]):
            pre = "%s_s%d_" % (name, si)
            bits_i, beta0, x, y, design, (bits_o, beta, T, nmse, expect), trace = _run_oracle(
                m["code_params"], m["decode_params"], m["awgn_var"], seed)
            assert np.array_equal(bits_i, g[pre + "bits_in"])
            assert np.array_equal(x, g[pre + "x"]) and np.array_equal(y, g[pre + "y"])
            assert T == int(g[pre + "t_final"]) and bool(expect) == bool(g[pre + "expect_err"])
            assert np.array_equal(bits_o, g[pre + "bits_out"])
            assert np.array_equal(beta.reshape(-1, design.M).argmax(1), g[pre + "idx"])
            assert np.array_equal(nmse, g[pre + "nmse"])
            assert np.array_equal(np.array([np.broadcast_to(t["tau"], (design.Lc,)) for t in trace]), g[pre + "tau_it"])
            assert np.array_equal(np.array([np.broadcast_to(t["psi"], (design.Lc,)) for t in trace]), g[pre + "psi_it"])
            assert np.array_equal(trace[-1]["s"].astype(np.float32), g[pre + "s_last"])


def test_implicit_operators_equal_closed_form_matrix():
    """Pins the third-party DCT at the boundary: Ab/Az == explicit
    sqrt(W/L) sqrt(2) cos(pi o0 (2 o1 + 1) / 2w)."""
    rng = np.random.RandomState(0)
    for W, L, M, n in ((np.array(3.0), 8, 8, 40), (np.array([1.0, 2.5]), 8, 4, 20),
                       (so.sc_basic(np.array(2.0), 2, 3), 6, 8, 32)):
        d = so.Design(W, L, M, n, [3, 4])
        A = d.dense()
        x, y = rng.randn(L * M), rng.randn(n)
        np.testing.assert_allclose(d.Ab(x), A @ x, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(d.Az(y), A.T @ y, rtol=1e-12, atol=1e-12)


def test_facade_host_logic_matches_oracle():
    """Base matrix, orderings and bit maps of the product façade == oracle."""
    cp = dict(P=15.0, R=1.4, L=64, M=16, spatially_coupled=True, omega=3, Lambda=8, power_allocated=True, B=2,
              R_PA_ratio=0.5)
    W1 = sparc.create_base_matrix(**dict(cp, awgn_var=1.0))
    W2 = so.create_base_matrix(**dict(cp, awgn_var=1.0))
    assert np.array_equal(W1, W2)
    n, _ = so.code_dims(cp, W2)
    d1, d2 = sparc.Design(W1, 64, 16, n, [7, 8]), so.Design(W2, 64, 16, n, [7, 8])
    assert np.array_equal(d1.order0, d2.order0) and np.array_equal(d1.order1, d2.order1) and d1.w == d2.w
    bits = sparc.rnd_bin_arr(64 * 4, [7, 8])
    assert np.array_equal(bits, so.rnd_bin_arr(64 * 4, [7, 8]))
    beta = sparc.bin_arr_2_msg_vector(bits, 16)
    assert np.array_equal(beta, so.bin_arr_2_msg_vector(bits, 16))
    assert np.array_equal(sparc.msg_vector_2_bin_arr(beta, 16), bits)     # reference sparc.py:1003-1008
    assert np.array_equal(sparc.msg_vector_map_estimator(beta + 0.1, 16), beta)


def test_code_params_semantics():
    """In-place normalisation, defaults and assertion-style validation (sparc.py:77-170)."""
    cp = dict(P=15.0, R=1.5, L=1024, M=32, junk=1)
    sparc.check_code_params(cp)
    assert "junk" not in cp and cp["power_allocated"] is False and cp["complex"] is False
    with pytest.raises(AssertionError):
        sparc.check_code_params(dict(P=15, R=1.5, L=1024, M=32))          # P must be float
    with pytest.raises(AssertionError):
        sparc.check_code_params(dict(P=15.0, R=1.5, L=1024, M=33))        # M power of 2
    with pytest.raises(Exception, match="Need code parameters"):
        sparc.check_code_params(dict(P=15.0, R=1.5, L=1024))
    with pytest.raises(AssertionError, match="B must divide L"):
        sparc.check_code_params(dict(P=15.0, R=1.5, L=1000, M=32, power_allocated=True, B=32, R_PA_ratio=0.9))
    dp = dict(t_max=25)
    sparc.check_decode_params(dp)
    assert dp == dict(t_max=25, rtol=1e-6, phi_est_method=1)
    with pytest.raises(Exception, match="Need decode parameters"):
        sparc.check_decode_params({})
    with pytest.raises(AssertionError, match="K=2"):                # a real design modulates with 2-PSK only
        sparc.sparc_encode(dict(P=15.0, R=1.5, L=64, M=8, modulated=True, K=4), 1.0, [1, 2])


@pytest.mark.reference
def test_oracle_bit_exact_vs_imported_reference():
    import ref_shims
    R = ref_shims.load()
    cp = dict(P=12.0, R=1.2, L=128, M=16, power_allocated=True, B=4, R_PA_ratio=0.8)
    for seed in ([11, 12], [13, 14]):
        cpr, dpr = dict(cp), dict(t_max=20)
        bi, b0, x, Ab, Az = R["sparc"].sparc_encode(cpr, 1.0, seed)
        y = R["sparc_sim"].awgn_channel(x, 1.0, seed)
        bo, beta, T, nmse, ex = R["sparc"].sparc_decode(y, cpr, dpr, 1.0, seed, b0, Ab, Az)
        bits_i, beta0, x2, y2, design, (bits_o, beta2, T2, nmse2, ex2), _ = _run_oracle(cp, dict(t_max=20), 1.0, seed)
        assert np.array_equal(x, x2) and np.array_equal(y, y2) and T == T2
        assert np.array_equal(bo, bits_o) and np.array_equal(beta, beta2) and np.array_equal(nmse, nmse2)


def test_oracle_matches_reference_golden_k2_modulated():
    """Real 2-PSK modulated SPARCs (sparc.py:433-441, 488-492): goldens of
    oracle/make_golden_sparc.py mod."""
    g = np.load(os.path.join(GOLDEN, "sparc_mod_reference.npz"))
    meta = json.loads(str(g["meta_json"]))
    assert set(meta) == {"k2_reg", "k2_pa_m64", "k2_sc_m16"}
    for name, m in meta.items():
        for si, seed in enumerate(m["seeds"]):
            pre = "%s_s%d_" % (name, si)
            bits_i, beta0, x, y, design, (bits_o, beta, T, nmse, expect), trace = _run_oracle(
                m["code_params"], m["decode_params"], m["awgn_var"], seed)
            assert np.array_equal(bits_i, g[pre + "bits_in"])
            assert np.array_equal(x, g[pre + "x"]) and np.array_equal(y, g[pre + "y"])
            assert T == int(g[pre + "t_final"]) and bool(expect) == bool(g[pre + "expect_err"])
            assert np.array_equal(bits_o, g[pre + "bits_out"])
            dense = beta.reshape(-1, design.M)
            idx = np.abs(dense).argmax(1)
            assert np.array_equal(idx, g[pre + "idx"]) and np.array_equal(dense[np.arange(len(idx)), idx], g[pre + "val"])
            assert np.array_equal(nmse, g[pre + "nmse"])
            assert np.array_equal(np.array([np.broadcast_to(t["tau"], (design.Lc,)) for t in trace]), g[pre + "tau_it"])
            assert np.array_equal(np.array([np.broadcast_to(t["psi"], (design.Lc,)) for t in trace]), g[pre + "psi_it"])
            assert np.array_equal(trace[-1]["s"].astype(np.float32), g[pre + "s_last"])


def test_oracle_matches_reference_golden_complex():
    """Complex designs (sub-sampled DFT, sparc.py:593-646) with K = 1, 2, 4 and 8: goldens of
    oracle/make_golden_sparc.py complex."""
    g = np.load(os.path.join(GOLDEN, "sparc_complex_reference.npz"))
    meta = json.loads(str(g["meta_json"]))
    assert set(meta) == {"c_k1", "c_k2", "c_k4_pa", "c_k8_sc_m16"}
    for name, m in meta.items():
        cp = so.normalise_code_params(m["code_params"])
        K = so.modulation(cp)
        for si, seed in enumerate(m["seeds"]):
            pre = "%s_s%d_" % (name, si)
            bits_i, beta0, x, y, design, (bits_o, beta, T, nmse, expect), trace = _run_oracle(
                cp, m["decode_params"], m["awgn_var"], seed)
            assert design.csparc and np.iscomplexobj(x)
            assert np.array_equal(bits_i, g[pre + "bits_in"])
            assert np.array_equal(x, g[pre + "x"]) and np.array_equal(y, g[pre + "y"])
            assert T == int(g[pre + "t_final"]) and bool(expect) == bool(g[pre + "expect_err"])
            assert np.array_equal(bits_o, g[pre + "bits_out"])
            dense = beta.reshape(-1, design.M)
            loc = np.abs(dense).argmax(1)
            assert np.array_equal(loc, g[pre + "sym"] % design.M)
            if K > 1:
                assert np.array_equal(dense[np.arange(len(loc)), loc], so.psk_constel(K)[g[pre + "sym"] // design.M])
            assert np.array_equal(nmse, g[pre + "nmse"])
            assert np.array_equal(np.array([np.broadcast_to(t["tau"], (design.Lc,)) for t in trace]), g[pre + "tau_it"])
            assert np.array_equal(np.array([np.broadcast_to(t["psi"], (design.Lc,)) for t in trace]), g[pre + "psi_it"])
            assert np.array_equal(trace[-1]["s"].astype(np.complex64), g[pre + "s_last"])


def test_complex_operators_equal_closed_form_matrix():
    """A = sqrt(W/L) exp(-2 pi i o0 o1 / w): Ab = A x, Az = A^H y."""
    rng = np.random.RandomState(0)
    for W, L, M, n in ((np.array(3.0), 8, 8, 20), (so.sc_basic(np.array(2.0), 2, 3), 6, 8, 32)):
        d = so.Design(W, L, M, n, [3, 4], csparc=True)
        A = d.dense()
        x = rng.randn(L * M) + 1j * rng.randn(L * M)
        y = rng.randn(n) + 1j * rng.randn(n)
        assert np.allclose(d.Ab(x), A @ x, atol=1e-11) and np.allclose(d.Az(y), A.conj().T @ y, atol=1e-11)
        assert not np.isin(d.order0[d.Wm != 0], [0, d.w // 2]).any()

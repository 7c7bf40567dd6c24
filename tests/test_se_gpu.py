"""State evolution on the GPU against the reference's golden trajectories (same seeded
Monte-Carlo draws) and the numpy oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
import sparc_se_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "se_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


@pytest.mark.parametrize("name", ["reg", "reg_m512_low_snr", "pa", "sc", "sc_wave", "k2", "k2_sc_pa", "c_k1", "c_k4_pa", "c_k8_sc"])
def test_sparc_se_matches_reference_golden(gold, name):
    from ldpc_sparc_b200 import sparc_se
    g, meta = gold
    m = meta[name]
    np.random.seed(m["seed"])
    psi, tau = sparc_se.sparc_se(m["awgn_var"], dict(m["code_params"]), m["t_max"], m["mc_samples"])
    assert psi.shape == g[name + "_psi"].shape
    assert np.allclose(psi, g[name + "_psi"], rtol=1e-9, atol=1e-12)       # fp64 kernel: summation order only
    assert np.allclose(tau, g[name + "_tau"], rtol=1e-9, atol=0)
    # fp32 samples: the expectation moves by the rounding of u, far below the Monte-Carlo error
    np.random.seed(m["seed"])
    psi32, _ = sparc_se.sparc_se(m["awgn_var"], dict(m["code_params"]), m["t_max"], m["mc_samples"], precision="fp32")
    assert np.allclose(psi32, g[name + "_psi"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("K", [1, 2])
def test_expectation_matches_oracle_and_survives_small_tau(K):
    from ldpc_sparc_b200 import sparc_se
    u = np.random.RandomState(5).randn(3001, 64)                           # ragged last CTA
    for tau in (2.0, 0.3, 0.05):
        assert abs(sparc_se.sparc_se_E(tau, K, u) - so.se_expectation(tau, K, u)) < 1e-12
    # 1 / tau > 709: the reference's exp overflows to nan, the shifted form gives the limit 1
    assert sparc_se.sparc_se_E(1e-3, K, u) == 1.0
    with pytest.raises(ValueError):
        sparc_se.sparc_se_E(0.3, 4, u)                                    # K > 2 needs complex samples


@pytest.mark.parametrize("K", [4, 8, 16])
def test_complex_expectation_matches_oracle(K):
    from ldpc_sparc_b200 import sparc_se
    rng = np.random.RandomState(6)
    u = rng.randn(1500, 32) + 1j * rng.randn(1500, 32)
    for tau in (2.0, 0.3, 0.05):
        assert abs(sparc_se.sparc_se_E(tau, K, u) - so.se_expectation(tau, K, u)) < 1e-12
    assert abs(sparc_se.sparc_se_E(0.3, K, u, precision="fp32") - so.se_expectation(0.3, K, u)) < 1e-5


def test_large_sample_count_runs_fast():
    import time
    from ldpc_sparc_b200 import sparc_se
    np.random.seed(1)
    t0 = time.time()
    psi, tau = sparc_se.sparc_se(1.0, dict(P=15.0, R=1.5, M=32, power_allocated=True, B=32, R_PA_ratio=0.9), 25, 200000)
    assert psi.shape == (25, 32) and np.all(psi[-1] < 1e-3)
    assert time.time() - t0 < 60

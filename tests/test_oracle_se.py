"""The state-evolution oracle against the golden vectors produced by the unmodified
reference (oracle/make_golden_se.py), and the host-side parameter check."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
import sparc_se_oracle as so


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "se_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


def test_golden_is_finite_and_covers_both_regimes(gold):
    g, meta = gold
    assert set(meta) >= {"reg", "pa", "sc", "sc_wave", "k2", "k2_sc_pa", "reg_m512_low_snr"}
    for name in meta:
        assert np.all(np.isfinite(g[name + "_psi"])) and np.all(np.isfinite(g[name + "_tau"]))
    assert g["reg_psi"][-1] < 1e-6 and g["reg_m512_low_snr_psi"][-1] > 0.5     # decodes / stalls


@pytest.mark.parametrize("name", ["reg", "reg_m512_low_snr", "pa", "sc", "sc_wave", "k2", "k2_sc_pa", "c_k1", "c_k4_pa", "c_k8_sc"])
def test_oracle_matches_reference(gold, name):
    g, meta = gold
    m = meta[name]
    cp = m["code_params"]
    np.random.seed(m["seed"])
    u = np.random.randn(m["mc_samples"], cp["M"])              # the draws sparc_se.py:150-153 make
    if cp.get("modulated") and cp["K"] > 2:
        u = u + 1j * np.random.randn(m["mc_samples"], cp["M"])
    psi, tau = so.state_evolution(m["awgn_var"], cp, m["t_max"], u)
    assert psi.shape == g[name + "_psi"].shape
    assert np.allclose(psi, g[name + "_psi"], rtol=1e-12, atol=1e-14)
    assert np.allclose(tau, g[name + "_tau"], rtol=1e-12, atol=0)


def test_check_code_params_se():
    from ldpc_sparc_b200 import sparc_se
    cp = dict(P=15.0, R=1.3, M=32, L=256, junk=1)
    sparc_se.check_code_params(cp)
    assert cp == dict(P=15.0, R=1.3, M=32, complex=False, modulated=False, power_allocated=False,
                      spatially_coupled=False)                                  # L and unknown keys are dropped
    with pytest.raises(Exception, match="Need code parameters"):
        sparc_se.check_code_params(dict(P=15.0, R=1.3))
    with pytest.raises(AssertionError):
        sparc_se.check_code_params(dict(P=15.0, R=1.3, M=48))
    with pytest.raises(AssertionError, match="K=2"):
        sparc_se.check_code_params(dict(P=15.0, R=1.3, M=32, modulated=True, K=4))
    with pytest.raises(Exception, match="Need code parameters"):
        sparc_se.check_code_params(dict(P=15.0, R=1.3, M=32, power_allocated=True, B=4))


@pytest.mark.reference
def test_check_code_params_se_against_reference():
    import ref_shims
    ref_shims.load()
    import sparc_se as ref_se
    from ldpc_sparc_b200 import sparc_se
    for cp in (dict(P=15.0, R=1.3, M=32, L=8), dict(P=15.0, R=1.4, M=16, modulated=True, K=2, spatially_coupled=True,
                                                   omega=3, Lambda=8, power_allocated=True, B=2, R_PA_ratio=0.5)):
        a, b = dict(cp), dict(cp)
        sparc_se.check_code_params(a)
        ref_se.check_code_params(b)
        assert a == b

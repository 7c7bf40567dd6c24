"""Host-side pieces of the simulation drivers (no GPU): parameter bookkeeping against the
values the reference's param_calc produces, error-rate helpers, channel model."""
import os
import sys

import numpy as np
import pytest

from ldpc_sparc_b200 import param_calc as pc
from ldpc_sparc_b200 import sparc_sim, sparc_sim_new


def test_param_calc_config5():
    # performance_plots_general.py:27-47: 802.16 r1/2 z=150, mults=4, logM=1, R_sparc_ldpc=1
    rate, L_sparc, L_sparc_ldpc, lengths = pc.param_calc(4, 1, '802.16', '1/2', 1 / 2, 150, 1)
    assert (rate, L_sparc, L_sparc_ldpc) == (0.5, 7200, 14400)
    assert lengths == {'k_ldpc': 7200, 'mults': 4, 'L_unprotected': 0}
    rate, L_sparc, L_sparc_ldpc, lengths = pc.param_calc(1, 1, '802.16', '1/2', 1 / 2, 150, 1)
    assert (rate, L_sparc, L_sparc_ldpc, lengths['k_ldpc']) == (0.5, 1800, 3600, 1800)


def test_param_calc_semi_protected():
    # performance_plots_general.py:50-64 defaults: R=0.8, mults=3, 100 % protected, M=64, 802.11n r5/6 z=81
    L_sparc, R_sl, L_sl, lengths, rate = pc.param_calc_semi_protected(0.8, 3, 1.0, 64, '802.11n', '5/6', 5 / 6, 81)
    assert (L_sparc, L_sl) == (810, 972)
    assert lengths == {'k_ldpc': 4860, 'mults': 3, 'L_unprotected': 0}
    assert abs(R_sl - 5832 / 6075) < 1e-12 and abs(rate - 4860 / 6075) < 1e-12
    # 60 % protected: unprotected bits rounded up to whole sections
    L_sparc, R_sl, L_sl, lengths, rate = pc.param_calc_semi_protected(0.8, 3, 0.6, 64, '802.11n', '5/6', 5 / 6, 81)
    assert lengths['L_unprotected'] == 540 and L_sl == 972 + 540


@pytest.mark.reference
@pytest.mark.parametrize("args", [(4, 1, '802.16', '1/2', 0.5, 150, 1), (2, 2, '802.11n', '1/2', 0.5, 81, 1.2),
                                  (3, 6, '802.11n', '5/6', 5 / 6, 81, 1.0)])
def test_param_calc_against_reference(args):
    import ref_shims
    ref_shims.load()
    sys.path.insert(0, "/root/reference/ldpc_sparc")
    import param_calc as ref_pc   # noqa: E402  (ldpc_sparc/param_calc.py)
    assert pc.param_calc(*args) == ref_pc.param_calc(*args)
    semi = (0.8, 3, 0.75, 64, '802.11n', '5/6', 5 / 6, 81)
    assert pc.param_calc_semi_protected(*semi) == ref_pc.param_calc_semi_protected(*semi)


def test_error_rates_and_channel():
    a = np.array([0, 1, 1, 0, 1, 0], dtype=bool)
    b = np.array([0, 1, 0, 0, 1, 1], dtype=bool)
    assert sparc_sim.calc_ber(a, b) == 2 / 6
    with pytest.raises(AssertionError):
        sparc_sim.calc_ber(a.astype(int), b)
    beta0 = np.zeros(12)
    beta0[[1, 4, 10]] = 1.0
    beta = np.zeros(12)
    beta[[1, 5, 10]] = 1.0
    ser, loc, num = sparc_sim.calc_ser(beta0, beta, 3)
    assert (ser, list(loc), num) == (1 / 3, [1], 1)
    x = np.linspace(-1, 1, 50)
    for mod in (sparc_sim, sparc_sim_new):
        y = mod.awgn_channel(x, 0.25, [3, 4])
        assert np.array_equal(y, x + 0.5 * np.random.RandomState([3, 4]).randn(50))
        with pytest.raises(AssertionError):
            mod.awgn_channel(x.reshape(5, 10), 0.25, 1)

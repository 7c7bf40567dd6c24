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


def test_modulated_bit_maps_and_error_rates():
    """Real 2-PSK helpers (sparc.py:203-400, sparc_sim.py:100-175): known answers."""
    from ldpc_sparc_b200 import sparc
    assert [sparc.bin2gray(i) for i in range(8)] == [0, 1, 3, 2, 6, 7, 5, 4]
    assert all(sparc.gray2bin(sparc.bin2gray(i)) == i for i in range(1 << 12))
    assert np.array_equal(sparc.psk_constel(2), [1, -1])
    assert np.allclose(sparc.psk_constel(8)[3], np.exp(2j * np.pi * 3 / 8))
    bits = np.array([1, 0, 1, 0, 0, 1, 1, 1], dtype=bool)            # M = 8, K = 2: (loc 5, +1), (loc 3, -1)
    v = sparc.bin_arr_2_msg_vector(bits, 8, 2)
    assert np.array_equal(np.flatnonzero(v), [5, 8 + 3]) and v[5] == 1 and v[11] == -1
    assert np.array_equal(sparc.msg_vector_2_bin_arr(v, 8, 2), bits)
    assert np.array_equal(sparc.section_indices(bits, 8, 2), [5, 8 + 3])
    assert np.array_equal(sparc.indices_2_bin_arr(np.array([5, 11]), 8, 2), bits)
    s = np.array([0.1, -3.0, 0.5, 2.0, 0.2, 0.1, 0.0, -0.1, 1.0, 0.3, 0.2, 0.1, 0.0, 0.0, 0.0, 0.9])
    assert np.array_equal(np.flatnonzero(sparc.msg_vector_map_estimator(s, 8, 2)), [1, 8])
    assert sparc.msg_vector_map_estimator(s, 8, 2)[1] == -1
    v4 = sparc.bin_arr_2_msg_vector(np.array([1, 0, 1, 1, 1, 0, 1, 1, 0, 1], dtype=bool), 8, 4)   # value bits 11 -> Gray index 2
    assert v4.dtype == complex and v4[5] == -1 and v4[8 + 3] == 1j                                 # 01 -> index 1 -> +i
    assert np.array_equal(sparc.msg_vector_2_symbols(v4, 8, 4), [5 + 8 * 2, 3 + 8 * 1])
    beta0 = np.zeros(12)
    beta0[[1, 4, 10]] = [1, -1, 1]
    beta = np.zeros(12)
    beta[[1, 5, 10]] = [1, -1, -1]
    rates, locs, nums = sparc_sim.calc_ler_ver(beta0, beta, 3, 2)
    assert nums == (1, 1, 2) and rates == (1 / 3, 1 / 3, 2 / 3)
    assert [list(x) for x in locs] == [[1], [2], [1, 2]]


@pytest.mark.reference
def test_modulated_helpers_against_reference():
    import ref_shims
    R = ref_shims.load()
    from ldpc_sparc_b200 import sparc
    ref = R["sparc"]
    for K in (2, 4, 16):
        bits = np.random.RandomState(K).randint(0, 2, int(np.log2(K)) * 9).astype(bool)
        assert np.allclose(sparc.psk_mod(bits, K), ref.psk_mod(bits, K))
        assert np.array_equal(sparc.psk_demod(sparc.psk_mod(bits, K), K), bits)
    for K in (1, 2):
        assert np.array_equal(sparc.rnd_msg_vector(32, 16, [3, 4], K), ref.rnd_msg_vector(32, 16, [3, 4], K))
        bits = np.random.RandomState(1).randint(0, 2, (4 + (K == 2)) * 32).astype(bool)
        v = sparc.bin_arr_2_msg_vector(bits, 16, K)
        assert np.array_equal(v, ref.bin_arr_2_msg_vector(bits, 16, K))
        assert np.array_equal(sparc.msg_vector_2_bin_arr(v, 16, K), ref.msg_vector_2_bin_arr(v, 16, K))
        s = np.random.RandomState(2).randn(32 * 16)
        assert np.array_equal(sparc.msg_vector_map_estimator(s, 16, K), ref.msg_vector_map_estimator(s, 16, K))
    b0, b1 = ref.rnd_msg_vector(64, 8, 5, 2), ref.rnd_msg_vector(64, 8, 6, 2)
    mine, theirs = sparc_sim.calc_ler_ver(b0, b1, 64, 2), R["sparc_sim"].calc_ler_ver(b0, b1, 64, 2)
    assert mine[0] == theirs[0] and mine[2] == theirs[2]
    assert all(np.array_equal(a, b) for a, b in zip(mine[1], theirs[1]))

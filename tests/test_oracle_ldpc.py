"""CPU: the C oracle against the reference's golden vectors and compiled reference."""
import os

import numpy as np
import pytest

import ldpc_oracle as orc
from conftest import ALL_CODES, GOLDEN
from ldpc_sparc_b200 import ldpc


def _graph(c):
    return orc.Graph(c.vdeg, c.cdeg, c.intrlv)


def test_oracle_matches_reference_golden_c2(golden_ldpc):
    """802.11n r1/2 z=81, 50 iterations, three Es/N0: app and it bit-exact vs the
    reference's c_ldpc.c outputs stored by oracle/make_golden.py."""
    c = ldpc.code("802.11n", "1/2", 81)
    g = _graph(c)
    for tag in ("c2_snr12", "c2_snr16", "c2_snr20"):
        llr, app_ref, it_ref = (golden_ldpc[tag + s] for s in ("_llr", "_app", "_it"))
        app, it = orc.decode_batch(g, llr, 50)
        assert np.array_equal(it, it_ref)
        assert np.array_equal(app, app_ref)   # same libm, same operation order


def test_oracle_matches_reference_golden_families(golden_ldpc):
    for i, name in enumerate(golden_ldpc["fam_names"]):
        std, rate, z, pt = str(name).split("|")
        c = ldpc.code(std, rate, int(z), pt)
        g = _graph(c)
        llr = golden_ldpc["fam%d_llr" % i]
        for algo in ("sumprod2", "sumprod"):
            for f in range(2):
                app, it = orc.decode(g, llr[f], 40, algo)
                assert it == golden_ldpc["fam%d_it_%s" % (i, algo)][f]
                np.testing.assert_array_equal(app, golden_ldpc["fam%d_app_%s" % (i, algo)][f])
        # the noiseless frame of the reference's test (test_ldpc.py:60-69)
        assert golden_ldpc["fam%d_it_sumprod2" % i][1] == 0


def test_oracle_lxor_lxfb_known_answers(golden_ldpc):
    for (a, b), (o0, o1) in zip(golden_ldpc["lxor_in"], golden_ldpc["lxor_out"]):
        assert orc.lxor(a, b, 0) == o0 and orc.lxor(a, b, 1) == o1
    for v, ret, out in zip(golden_ldpc["lxfb_in"], golden_ldpc["lxfb_ret"], golden_ldpc["lxfb_out"]):
        r, o = orc.lxfb(v, 1)
        assert r == ret
        np.testing.assert_array_equal(o, out)


def test_graph_layout_fixture():
    """ldpc_jossy/src/ldpc802.16.81.h == code('802.16','1/2',81)."""
    fx = np.load(os.path.join(GOLDEN, "ldpc_802_16_z81_graph.npz"))
    c = ldpc.code("802.16", "1/2", 81)
    assert (c.Nv, c.Nc, c.Nmsg) == (int(fx["Nv"]), int(fx["Nc"]), int(fx["Nmsg"]))
    assert np.array_equal(c.intrlv, fx["intrlv"])
    assert np.array_equal(c.vdeg, fx["vdeg"]) and np.array_equal(c.cdeg, fx["cdeg"])


@pytest.mark.parametrize("standard,rate,z,ptype", ALL_CODES)
def test_oracle_noiseless_all_codes(standard, rate, z, ptype):
    """Reference test_ldpc.py semantics on the oracle: graph sums, x.H^T = 0 and
    it == 0 with zero bit errors for y = 10(0.5-x), both decoders."""
    c = ldpc.code(standard, rate, z, ptype)
    assert len(c.proto[0]) == 24
    assert np.sum(c.vdeg) == np.sum(c.cdeg) == len(c.intrlv)
    H = c.pcmat()
    assert np.sum(c.vdeg) == np.sum(H)
    g = _graph(c)
    rng = np.random.RandomState(z)
    u = rng.randint(0, 2, (4, c.K))
    x = c.encode(u)
    assert np.count_nonzero(x.dot(H.T) % 2) == 0
    for f in range(4):
        y = 10 * (0.5 - x[f])
        for algo in ("sumprod", "sumprod2"):
            app, it = orc.decode(g, y, 200, algo)
            assert it == 0
            assert np.count_nonzero((app < 0) != x[f]) == 0


@pytest.mark.skipif(orc.ref_lib() is None, reason="oracle/_ref/c_ldpc_ref.so not built")
def test_oracle_bit_exact_vs_compiled_reference():
    """Random noisy frames: restatement == reference binary, bit for bit."""
    rng = np.random.RandomState(3)
    for fam, snr in ((("802.11n", "1/2", 81, "A"), 1.4), (("802.16", "3/4", 54, "B"), 4.0),
                     (("802.16", "5/6", 27, "A"), 5.0)):
        c = ldpc.code(*fam)
        g = _graph(c)
        x = c.encode(rng.randint(0, 2, (5, c.K)))
        s2 = 10 ** (-snr / 10)
        llr = 2 / s2 * ((1 - 2.0 * x) + np.sqrt(s2) * rng.randn(*x.shape))
        for algo in ("sumprod2", "sumprod"):
            for f in range(len(llr)):
                a1, i1 = orc.decode(g, llr[f], 30, algo, impl="oracle")
                a2, i2 = orc.decode(g, llr[f], 30, algo, impl="ref")
                assert i1 == i2
                np.testing.assert_array_equal(a1, a2)

"""Simulation drivers on the GPU: the single-frame mirrors reproduce the reference's own
end-to-end runs (golden vectors made by oracle/make_golden_sparc.py from the unmodified
reference), the batched Monte-Carlo forms are deterministic, count what they should and
land on the error rates of the per-frame path."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "sparc_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name", ["reg", "sc_pa_m16", "reg_m64"])
def test_sparc_sim_reproduces_reference_run(gold, name, precision):
    from ldpc_sparc_b200 import sparc_sim
    g, meta = gold
    m = meta[name]
    for si, seed in enumerate(m["seeds"]):
        pre = "%s_s%d_" % (name, si)
        res = sparc_sim.sparc_sim(dict(m["code_params"]), dict(m["decode_params"]), m["awgn_var"], seed,
                                  precision=precision)
        bits_in, bits_out = g[pre + "bits_in"], g[pre + "bits_out"]
        assert res['ber'] == np.count_nonzero(bits_in ^ bits_out) / bits_in.size        # bit-exact decisions
        assert res['t_final'] == int(g[pre + "t_final"])
        logM = int(np.log2(m["code_params"]["M"]))
        truth = bits_in.reshape(-1, logM).astype(int).dot(1 << np.arange(logM)[::-1])
        assert res['num_of_sec_errs'] == int((truth != g[pre + "idx"]).sum())
        assert res['cer'] == float(res['ber'] > 0)
        assert res['detect'] == float(not ((res['ber'] > 0) ^ bool(g[pre + "expect_err"])))
        if precision == "fp64":
            assert np.allclose(res["nmse"], g[pre + "nmse"], rtol=1e-5, atol=1e-7)     # nmse out is fp32
        else:                                                                          # fp32 state: 1e-4 relative
            assert np.allclose(res["nmse"], g[pre + "nmse"], rtol=1e-4, atol=1e-6)


def test_sparc_sim_batch_counts_and_determinism():
    from ldpc_sparc_b200 import sparc_sim
    cp = dict(P=15.0, R=1.3, L=256, M=32)
    dp = dict(t_max=25)
    a = sparc_sim.sparc_sim_batch(cp, dp, 1.0, 640, rand_seed=(1, 2), seed=3, chunk=256)
    b = sparc_sim.sparc_sim_batch(cp, dp, 1.0, 640, rand_seed=(1, 2), seed=3, chunk=256)
    assert a == b                                             # same seeds -> same counters
    assert a['frames'] == 640 and a['n'] == int(round(256 * 5 / 1.3))
    assert a['section_errors'] <= a['bit_errors'] <= 5 * a['section_errors']
    assert a['frame_errors'] <= a['section_errors']
    assert 3 <= a['t_final'] <= 24
    # this operating point decodes essentially error-free in the reference (golden 'reg': no errors)
    assert a['ser'] < 2e-3 and a['detect'] > 0.95
    # low SNR: errors appear and the fp64 kernel agrees with the fp32 one on the error rate
    lo32 = sparc_sim.sparc_sim_batch(cp, dp, 2.2, 256, seed=1, chunk=256)
    lo64 = sparc_sim.sparc_sim_batch(cp, dp, 2.2, 256, seed=1, chunk=256, precision="fp64")
    assert lo64['ser'] > 0.01
    assert abs(lo32['ser'] - lo64['ser']) <= 0.02 * lo64['ser'] + 1e-4


def test_design_ensemble_pool():
    """n_designs: the reference draws a new design per trial (sparc_sim.py:15-24, performance_plots_general.py:101-118);
    a pool of D designs per batch averages over designs as well.  D = 1 reproduces the fixed-design counters; with D > 1
    every block of frames runs on its own design (first block = the fixed design's first block) and the totals stay
    deterministic; a block decoded on design d equals a single-design run on rand_seed + d."""
    from ldpc_sparc_b200 import sparc_sim, sparc_sim_new
    cp = dict(P=15.0, R=1.3, L=256, M=32)
    dp = dict(t_max=25)
    fixed = sparc_sim.sparc_sim_batch(cp, dp, 2.0, 768, rand_seed=(1, 2), seed=3, chunk=256)
    d1 = sparc_sim.sparc_sim_batch(cp, dp, 2.0, 768, rand_seed=(1, 2), seed=3, chunk=256, n_designs=1)
    assert {k: v for k, v in d1.items() if k != 'n_designs'} == {k: v for k, v in fixed.items() if k != 'n_designs'}
    d3 = sparc_sim.sparc_sim_batch(cp, dp, 2.0, 768, rand_seed=(1, 2), seed=3, chunk=256, n_designs=3)
    assert d3 == sparc_sim.sparc_sim_batch(cp, dp, 2.0, 768, rand_seed=(1, 2), seed=3, chunk=256, n_designs=3)
    assert d3['frames'] == 768 and d3['n_designs'] == 3
    assert d3['section_errors'] != fixed['section_errors'] and abs(d3['ser'] - fixed['ser']) < 0.5 * fixed['ser'] + 1e-3
    # block 1 of the pool run is block 1 of a fixed-design run on the second design: same messages, noise and design
    only_b = sparc_sim.sparc_sim_batch(cp, dp, 2.0, 512, rand_seed=(2, 3), seed=3, chunk=256)        # blocks 0, 1 on design 1
    two = sparc_sim.sparc_sim_batch(cp, dp, 2.0, 512, rand_seed=(1, 2), seed=3, chunk=256, n_designs=2)  # block 1 on design 1
    first = sparc_sim.sparc_sim_batch(cp, dp, 2.0, 256, rand_seed=(1, 2), seed=3, chunk=256)          # block 0 on design 0
    first_b = sparc_sim.sparc_sim_batch(cp, dp, 2.0, 256, rand_seed=(2, 3), seed=3, chunk=256)        # block 0 on design 1
    assert two['section_errors'] == first['section_errors'] + (only_b['section_errors'] - first_b['section_errors'])
    # dense-design sweep: pool of 2 designs, deterministic, D = 1 equals the default
    sp, spl, lp, lengths = _config5_small()
    kw = dict(sims=('SPARC+LDPC',), chunk=16, seed=5, precision="tc")
    a = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': 25}, [2.0], 32, **kw)
    b = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': 25}, [2.0], 32, n_designs=1, **kw)
    assert np.array_equal(a['bit_errors'], b['bit_errors'])
    c2 = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': 25}, [2.0], 32, n_designs=2, **kw)
    c3 = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': 25}, [2.0], 32, n_designs=2, **kw)
    assert np.array_equal(c2['bit_errors'], c3['bit_errors']) and np.all(c2['bits'] == a['bits'])


def _config5_small():
    from ldpc_sparc_b200.param_calc import param_calc
    mults, logM, standard, rate, int_rate, z = 1, 1, '802.16', '1/2', 1 / 2, 150
    overall, L_sparc, L_sl, lengths = param_calc(mults, logM, standard, rate, int_rate, z, 1)
    sp = {'P': 19.44, 'R': overall, 'L': L_sparc, 'M': 2}
    spl = {'P': 19.44, 'R': 1, 'L': L_sl, 'M': 2}
    lp = {'standard': standard, 'rate': rate, 'z': z, 'int_rate': int_rate, 'mults': mults}
    return sp, spl, lp, lengths


def test_single_frame_sims_run():
    from ldpc_sparc_b200 import sparc_sim_new
    sp, spl, lp, lengths = _config5_small()
    dp = {'t_max': 25}
    for fn, params, ldpc_bool in ((sparc_sim_new.sparc_ldpc_sim, sp, False), (sparc_sim_new.sparc_ldpc_sim, spl, True),
                                  (sparc_sim_new.sparc_ldpc_naive_sim, spl, True),
                                  (sparc_sim_new.sparc_ldpc_integrated_sim, spl, True)):
        bits_i, bits_o, ber = fn(params, lp, lengths, ldpc_bool, dp, 19.44 / 6.0, [11, 12])
        assert bits_i.shape == bits_o.shape == (1800,)
        assert ber == np.count_nonzero(np.asarray(bits_i) != np.asarray(bits_o)) / 1800
        assert ber < 0.02


@pytest.mark.parametrize("precision", ["tc", "fp64"])
def test_performance_sweep_small(tmp_path, precision):
    from ldpc_sparc_b200 import sparc_sim_new
    sp, spl, lp, lengths = _config5_small()
    path = str(tmp_path / "Test_0.npz")
    out = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': 25}, [2.0, 6.0], 48, chunk=32, seed=5,
                                          precision=precision, save_path=path)
    avg = out['ber_store_averages']
    assert avg.shape == (4, 2) and np.all(avg >= 0) and np.all(avg < 0.5)
    assert np.all(out['bits'] == 48 * 1800)
    assert np.all(out['ber_store_min'] <= avg + 1e-12) and np.all(avg <= out['ber_store_max'] + 1e-12)
    assert avg[0, 0] > 1e-3                       # uncoded-rate SPARC at snr = 2 makes errors
    assert np.all(avg[:, 1] <= avg[:, 0] + 1e-12)  # every scheme improves with SNR
    assert np.all(avg[1:, 1] == 0)                # LDPC-protected schemes are clean at snr = 6
    saved = np.load(path)
    assert set(saved.files) == {'ber_store_averages', 'ber_store_max', 'ber_store_min', 'snr_store'}
    assert np.array_equal(saved['ber_store_averages'], avg)


def test_config1_error_statistics_match_oracle():
    """Full-size statistical parity at BASELINE config 1: frame error rate, section error rate and
    mean iteration count of the batched GPU path over 4096 fresh frames against the oracle's
    statistics over 1536 frames (tests/golden/c1_statistics.json, oracle/make_golden_stats.py).
    Different frames on the two sides, so the bounds are 4 sigma of the two binomial / sample errors."""
    from ldpc_sparc_b200 import sparc_sim
    ref = json.load(open(os.path.join(GOLDEN, "c1_statistics.json")))
    n_ref, L = ref["frames"], ref["code_params"]["L"]
    fer_ref = ref["frames_in_error"] / n_ref
    ser_ref = ref["section_errors"] / (n_ref * L)
    F = 4096
    for precision in ("fp64", "fp32"):
        out = sparc_sim.sparc_sim_batch(dict(ref["code_params"]), dict(t_max=ref["t_max"]), ref["awgn_var"], F,
                                        rand_seed=tuple(ref["design_seed"]), seed=11, chunk=1024, precision=precision)
        sig_fer = np.sqrt(fer_ref * (1 - fer_ref) * (1 / n_ref + 1 / F))
        assert abs(out["cer"] - fer_ref) < 4 * sig_fer, (precision, out["cer"], fer_ref)
        sig_ser = np.sqrt(ref["section_errors_per_frame_var"] * (1 / n_ref + 1 / F)) / L
        assert abs(out["ser"] - ser_ref) < 4 * sig_ser, (precision, out["ser"], ser_ref)
        if precision == "fp64":          # fp32 meets the stop rule a few iterations later
            sig_t = ref["std_t_final"] * np.sqrt(1 / n_ref + 1 / F)
            assert abs(out["t_final"] - ref["mean_t_final"]) < 4 * sig_t, (out["t_final"], ref["mean_t_final"])
        else:
            assert ref["mean_t_final"] - 0.5 < out["t_final"] <= ref["t_max"] - 1


def test_campaign_cli(tmp_path, capsys):
    """python -m ldpc_sparc_b200.campaign: the three campaigns on tiny budgets."""
    from ldpc_sparc_b200 import campaign
    res = str(tmp_path / "results.txt")
    campaign.main(["ldpc", "--z", "27", "--points", "2", "--min-errors", "20", "--max-blocks", "4000", "--max-it", "30",
                   "--results", res])
    lines = open(res).read().strip().splitlines()
    assert len(lines) == 2 and "np." not in lines[0]
    std, rate, z, snr, nblocks, nerr, nbits, nbiterr, nit = eval(lines[0])
    assert (std, rate, z) == ("802.11n", "1/2", 27) and nerr == 20 and nbits == nblocks * 324
    capsys.readouterr()
    campaign.main(["sparc", "--L", "128", "--M", "16", "--R", "1.2", "--frames", "256", "--pa", "4", "0.9"])
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert out["frames"] == 256 and 0 <= out["ser"] < 0.05
    campaign.main(["sparc", "--L", "128", "--M", "16", "--R", "2.0", "--frames", "128", "--K", "4", "--complex"])
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert out["frames"] == 128 and out["n"] == int(round(128 * 6 / 2.0))
    npz = str(tmp_path / "Test_0.npz")
    campaign.main(["ldpc_sparc", "--mults", "1", "--frames", "32", "--snr", "2", "6", "2", "--npz", npz])
    saved = np.load(npz)
    assert saved["ber_store_averages"].shape == (4, 2) and np.all(saved["ber_store_averages"][1:, 1] == 0)


def test_sparc_sim_batch_on_the_tensor_core_path():
    """precision='tc' in the Monte-Carlo driver (explicit design, tcgen05 products): a spatially coupled code, same
    messages and noise as the fp32 FFT run -- identical section / frame error counters, deterministic."""
    from ldpc_sparc_b200 import sparc_sim
    cp = dict(P=15.0, R=1.4, L=256, M=32, spatially_coupled=True, omega=2, Lambda=8)
    dp = dict(t_max=25)
    a = sparc_sim.sparc_sim_batch(dict(cp), dp, 1.6, 384, rand_seed=(1, 2), seed=4, chunk=192, precision="tc")
    b = sparc_sim.sparc_sim_batch(dict(cp), dp, 1.6, 384, rand_seed=(1, 2), seed=4, chunk=192, precision="tc")
    f = sparc_sim.sparc_sim_batch(dict(cp), dp, 1.6, 384, rand_seed=(1, 2), seed=4, chunk=192, precision="fp32")
    assert a == b
    assert a['frames'] == 384 and a['section_errors'] > 0
    assert a['section_errors'] == f['section_errors'] and a['frame_errors'] == f['frame_errors']

"""GPU parity tests of the dense-design LDPC-SPARC path (sparc_sophie): AMP, the
section<->bit glue and the three AMP+BP decoders against the reference's goldens and
the numpy oracle."""
import json
import os

import numpy as np
import pytest

import sophie_oracle as sph
from conftest import GOLDEN
from ldpc_sparc_b200 import _lib, ldpc, sparc_new
from test_oracle_sophie import setup_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "sophie_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def test_dense_amp_matches_golden(gold):
    g, meta = gold
    for name, m in meta.items():
        pre, c, sp, dp, y, A = setup_case(g, m, name, 0)
        beta, s, tau = sparc_new.sparc_amp(y, sp, dp, A, precision="fp64", return_trace=True)
        assert rel(tau, g[pre + "amp_tau_sqr"]) < 1e-10
        assert rel(beta, g[pre + "amp_beta"]) < 1e-8 and rel(s, g[pre + "amp_s"]) < 1e-10
        beta32, s32, tau32 = sparc_new.sparc_amp(y, sp, dp, A, precision="fp32", return_trace=True)
        assert np.max(np.abs(tau32 - g[pre + "amp_tau_sqr"]) / g[pre + "amp_tau_sqr"]) < 1e-4      # north_star tolerance
        assert rel(beta32, g[pre + "amp_beta"]) < 1e-4
        assert np.array_equal(beta32.reshape(-1, sp["M"]).argmax(1), g[pre + "amp_beta"].reshape(-1, sp["M"]).argmax(1))


def test_glue_kernels_vs_oracle():
    rng = np.random.RandomState(0)
    for M in (2, 4, 8, 64):
        L, sq = 40, 1.7
        logM = int(np.log2(M))
        beta = rng.dirichlet(np.ones(M) * 0.3, size=L).ravel() * sq
        p_ref = sph.beta_estimate_to_bp_probs(beta, L, M, sq)
        p = sparc_new.beta_estimate_to_bp_probs(beta, L, M, sq)
        assert np.max(np.abs(p - p_ref)) < 1e-14
        probs = rng.uniform(0.01, 0.99, L * logM)
        b_ref = sph.bp_output_to_beta_estimate(probs, L, M, sq)
        b = sparc_new.bp_output_to_beta_estimate(probs, L, M, sq)
        assert np.max(np.abs(b - b_ref)) < 1e-13 * sq
        s = rng.randn(L * M)
        e_ref = sph.mmse_estimator(s, 0.37, 50, sq * sq / 50, M)
        e = sparc_new.msg_vector_mmse_estimator(s, 0.37, 50, sq * sq / 50, M)
        assert np.max(np.abs(e - e_ref)) < 1e-13
        e32 = sparc_new.msg_vector_mmse_estimator(s, 0.37, 50, sq * sq / 50, M, precision="fp32")
        assert np.max(np.abs(e32 - e_ref)) < 1e-5


def test_section_to_bit_saturated_fp32_sign():
    """fp32 section mass that sums to >= 1 (or <= 0) on one side of a bit must give the reference's clipped
    LLR +-log((1-1e-15)/1e-15) with the right sign, never a NaN turned into -lim (sparc_new.py:1166-1169)."""
    import torch
    from ldpc_sparc_b200 import _lib
    M, logM, L = 8, 3, 64
    sq = np.float32(1.7)
    rng = np.random.RandomState(5)
    beta = np.zeros((L, M), dtype=np.float32)
    for sec in range(L):
        # all of the mass on entries whose MSB is 0, split over 2-3 entries so the fp32 sum of beta_j/sq can exceed 1
        js = rng.choice(4, size=rng.randint(2, 4), replace=False)
        wts = rng.dirichlet(np.ones(len(js))).astype(np.float32)
        beta[sec, js] = np.nextafter(wts * sq, np.float32(np.inf))
    bd = torch.from_numpy(beta.reshape(1, -1)).cuda()
    probs = torch.empty((1, L * logM), dtype=torch.float32, device="cuda")
    llr = torch.empty_like(probs)
    _lib.check(_lib.lib().lsb_section_to_bit_dev_f32(bd.data_ptr(), probs.data_ptr(), llr.data_ptr(), L, M, float(sq),
                                                     torch.cuda.current_stream().cuda_stream), "section_to_bit")
    p = probs.cpu().numpy().reshape(L, logM)
    v = llr.cpu().numpy().reshape(L, logM)
    assert np.isfinite(v).all()
    assert (p[:, 0] >= 1.0).any(), "the case must exercise a probability that rounds to >= 1"
    lim = np.float32(34.538776)
    assert np.all(v[:, 0] > 30.0), "MSB is certainly 0: LLR must saturate positive"
    assert np.all(v[p >= 1.0] == lim) and np.all(v[p <= 0.0] == -lim)
    ref = np.log(np.clip(p.astype(np.float64), 1e-15, 1 - 1e-15)) - np.log1p(-np.clip(p.astype(np.float64), 1e-15, 1 - 1e-15))
    inside = (p > 0) & (p < 1)
    assert np.max(np.abs(v[inside] - ref[inside]) / np.maximum(1.0, np.abs(ref[inside]))) < 1e-5


def test_ldpc_bp_wrapper_vs_oracle():
    c = ldpc.code("802.16", "1/2", 27)
    rng = np.random.RandomState(2)
    x = np.concatenate([c.encode(rng.randint(0, 2, c.K)) for _ in range(2)])
    probs = np.clip(np.where(x == 0, 0.8, 0.2) + 0.25 * rng.randn(x.size), 0.0, 1.0)
    p_ref, _ = sph.ldpc_bp(probs, sph.Code(c), 6, False)
    p, _ = sparc_new.ldpc_bp(probs, c, 6, False)
    assert np.allclose(p, p_ref, rtol=1e-9, atol=1e-12)
    _, b_ref = sph.ldpc_bp(probs, sph.Code(c), 200, True)
    _, b = sparc_new.ldpc_bp(probs, c, 200, True)
    assert np.array_equal(b, b_ref) and b.shape == (2 * c.K,)


@pytest.mark.parametrize("name", ["z27_m2", "z27_m4", "n11_m8", "z54_m2_hi"])
def test_decoders_fp64_match_reference_golden(gold, name):
    g, meta = gold
    m = meta[name]
    for si in range(len(m["seeds"])):
        pre, c, sp, dp, y, A = setup_case(g, m, name, si)
        lp = dict(standard=m["standard"], rate=m["rate"], z=m["z"])
        lengths = dict(k_ldpc=c.K * m["mults"], mults=m["mults"], L_unprotected=0)
        bits = sparc_new.sparc_ldpc_decode(y, sp, lp, dp, True, lengths, A, precision="fp64")
        assert np.array_equal(bits, g[pre + "bits_separate"])
        bits, tau = sparc_new.naively_integrated_decoder(y, sp, lp, dp, A, precision="fp64", return_trace=True)
        assert rel(tau, g[pre + "naive_tau_sqr"]) < 1e-8
        assert np.array_equal(bits, g[pre + "bits_naive"])
        bits, tau, ons = sparc_new.integrated_decoder(y, sp, lp, dp, A, precision="fp64", return_trace=True)
        assert rel(tau, g[pre + "int_tau_sqr"]) < 1e-8
        assert rel(ons, g[pre + "int_onsager_sum"]) < 1e-8
        assert np.array_equal(bits, g[pre + "bits_integrated"])


def test_decoders_fp32_converging_case(gold):
    """fp32 mode on the operating point where the decoders converge (BER 0 in the
    reference): identical bits, tau^2 within 1e-4."""
    g, meta = gold
    m = meta["z54_m2_hi"]
    for si in range(2):
        pre, c, sp, dp, y, A = setup_case(g, m, "z54_m2_hi", si)
        lp = dict(standard=m["standard"], rate=m["rate"], z=m["z"])
        lengths = dict(k_ldpc=c.K * m["mults"], mults=m["mults"], L_unprotected=0)
        assert np.array_equal(sparc_new.sparc_ldpc_decode(y, sp, lp, dp, True, lengths, A, precision="fp32"),
                              g[pre + "bits_separate"])
        bits, tau = sparc_new.naively_integrated_decoder(y, sp, lp, dp, A, precision="fp32", return_trace=True)
        assert np.array_equal(bits, g[pre + "bits_naive"])
        assert np.max(np.abs(tau - g[pre + "naive_tau_sqr"]) / g[pre + "naive_tau_sqr"]) < 1e-4
        bits, tau, ons = sparc_new.integrated_decoder(y, sp, lp, dp, A, precision="fp32", return_trace=True)
        assert np.array_equal(bits, g[pre + "bits_integrated"])
        assert np.max(np.abs(tau - g[pre + "int_tau_sqr"]) / g[pre + "int_tau_sqr"]) < 1e-4
        # Onsager sums (closed-form fp32 kernel) against the reference's nested loops
        ref_ons = g[pre + "int_onsager_sum"]
        assert np.max(np.abs(ons - ref_ons)) <= 1e-4 * np.max(np.abs(ref_ons))


def test_encode_and_batched_decode(gold):
    """sparc_ldpc_encode reproduces the reference's codeword; a batch of frames sharing A
    decodes to the same bits as frame-by-frame calls."""
    g, meta = gold
    m = meta["z54_m2_hi"]
    pre, c, sp, dp, y, A = setup_case(g, m, "z54_m2_hi", 0)
    lp = dict(standard=m["standard"], rate=m["rate"], z=m["z"])
    lengths = dict(k_ldpc=c.K * m["mults"], mults=m["mults"], L_unprotected=0)
    ub, tb, beta0, x, A2 = sparc_new.sparc_ldpc_encode(sp, lp, lengths, True, m["seeds"][0])
    assert np.array_equal(ub, g[pre + "user_bits"]) and np.array_equal(tb, g[pre + "total_bits"].astype(bool))
    assert rel(x, g[pre + "x"]) < 1e-12 and np.array_equal(A2, A)
    rng = np.random.RandomState(1)
    ys = np.stack([g[pre + "x"] + np.sqrt(sp["P"] / 5.0) * rng.randn(len(y)) for _ in range(20)])
    for dec in (lambda yy, pr: sparc_new.sparc_ldpc_decode(yy, sp, lp, dp, True, lengths, A, precision=pr),
                lambda yy, pr: sparc_new.naively_integrated_decoder(yy, sp, lp, dp, A, precision=pr),
                lambda yy, pr: sparc_new.integrated_decoder(yy, sp, lp, dp, A, precision=pr)):
        batch = dec(ys, "fp64")
        assert batch.shape == (20, c.K * m["mults"])
        for f in (0, 7, 19):
            assert np.array_equal(batch[f], dec(ys[f], "fp64"))
        assert (batch != g[pre + "user_bits"]).mean() < 0.05
        assert (dec(ys, "fp32") != batch).mean() < 0.01


def test_tensor_core_products_vs_fp64():
    """tcgen05 split-fp16 products against float64 numpy: ~16 mantissa bits."""
    import ctypes as ct
    import torch
    from ldpc_sparc_b200 import _lib
    rng = np.random.RandomState(3)
    for n, ML, F in ((256, 512, 128), (384, 1160, 200), (1000, 2408, 37)):
        A = rng.randn(n, ML) / np.sqrt(n)
        eng = sparc_new._Engine(A, "tc")
        x = rng.randn(F, ML) * np.array([1.0, 1e-3, 30.0])[rng.randint(0, 3, (F, 1))]
        z = rng.randn(F, n)
        ab = eng.ab(eng.to_dev(x)).cpu().numpy().astype(np.float64)
        az = eng.az(eng.to_dev(z)).cpu().numpy().astype(np.float64)
        ref_ab, ref_az = x @ A.T, z @ A
        err_ab = np.max(np.abs(ab - ref_ab)) / np.max(np.abs(ref_ab))
        err_az = np.max(np.abs(az - ref_az)) / np.max(np.abs(ref_az))
        assert err_ab < 5e-5 and err_az < 5e-5, (n, ML, F, err_ab, err_az)
        # row-wise: every frame within 1e-4 of its own scale
        assert np.all(np.max(np.abs(ab - ref_ab), 1) < 1e-4 * np.max(np.abs(ref_ab), 1))


def test_tensor_core_mode_decoders(gold):
    """precision='tc' (tensor-core products, fp32 elsewhere) on the converging case: same
    bits as the reference, tau^2 within 1e-4."""
    g, meta = gold
    m = meta["z54_m2_hi"]
    pre, c, sp, dp, y, A = setup_case(g, m, "z54_m2_hi", 0)
    lp = dict(standard=m["standard"], rate=m["rate"], z=m["z"])
    lengths = dict(k_ldpc=c.K * m["mults"], mults=m["mults"], L_unprotected=0)
    beta, s, tau = sparc_new.sparc_amp(y, sp, dp, A, precision="tc", return_trace=True)
    assert np.max(np.abs(tau - g[pre + "amp_tau_sqr"]) / g[pre + "amp_tau_sqr"]) < 1e-4
    assert rel(beta, g[pre + "amp_beta"]) < 1e-4
    assert np.array_equal(sparc_new.sparc_ldpc_decode(y, sp, lp, dp, True, lengths, A, precision="tc"), g[pre + "bits_separate"])
    assert np.array_equal(sparc_new.naively_integrated_decoder(y, sp, lp, dp, A, precision="tc"), g[pre + "bits_naive"])
    assert np.array_equal(sparc_new.integrated_decoder(y, sp, lp, dp, A, precision="tc"), g[pre + "bits_integrated"])


def test_config5_small_variant_vs_oracle():
    """BASELINE config 5, parity variant of SURVEY section 8: 802.16 r1/2 z=150, mults=1,
    logM=1 (M=2), R=1 -> L = n = 3600, A = 3600 x 7200, P = 19.44, t_max = 25, at the two
    ends of the reference's SNR sweep (performance_plots_general.py:21-47,96-98)."""
    c = ldpc.code("802.16", "1/2", 150)
    oc = sph.Code(c)
    L, M, n = 3600, 2, 3600
    sp = dict(P=19.44, R=1.0, L=L, M=M)
    lp = dict(standard="802.16", rate="1/2", z=150)
    lengths = dict(k_ldpc=c.K, mults=1, L_unprotected=0)
    dp = dict(t_max=25)
    ub, tb, beta0, x, A = sph.encode(sp, oc, 1, 11, c.encode)
    rng = np.random.RandomState(11)
    for snr in (6.0, 3.0):
        y = x + np.sqrt(sp["P"] / snr) * rng.randn(n)
        ref_sep = sph.sparc_ldpc_decode(y, sp, oc, dp, True, lengths, A)
        ref_int = sph.integrated_decoder(y, sp, oc, dp, A)
        got_sep = sparc_new.sparc_ldpc_decode(y, sp, lp, dp, True, lengths, A, precision="fp64")
        got_int = sparc_new.integrated_decoder(y, sp, lp, dp, A, precision="fp64")
        assert np.array_equal(got_sep, ref_sep) and np.array_equal(got_int, ref_int), snr
        if snr == 6.0:
            assert np.array_equal(ref_int, ub)                       # decodes cleanly at the top of the sweep
            assert np.array_equal(sparc_new.integrated_decoder(y, sp, lp, dp, A, precision="tc"), ub)
            assert np.array_equal(sparc_new.naively_integrated_decoder(y, sp, lp, dp, A, precision="tc"),
                                  sph.naively_integrated_decoder(y, sp, oc, dp, A))


def test_config5_full_size_vs_oracle():
    """BASELINE config 5 at FULL size (performance_plots_general.py:21-47): 802.16 r1/2 z=150, mults=4, M=2, R=1 ->
    L = n = 14 400, A = 14 400 x 28 800 (3.3 GB in fp64), P = 19.44, t_max = 25, four LDPC blocks per frame.  The
    tensor-core mode -- the one the bench times -- against the numpy oracle on the same A and y: identical user bits
    for the separate and the integrated decoder, tau^2 trace of the plain AMP within 1e-4."""
    c = ldpc.code("802.16", "1/2", 150)
    oc = sph.Code(c)
    mults, L, M, n = 4, 14400, 2, 14400
    sp = dict(P=19.44, R=1.0, L=L, M=M)
    lp = dict(standard="802.16", rate="1/2", z=150)
    lengths = dict(k_ldpc=c.K * mults, mults=mults, L_unprotected=0)
    dp = dict(t_max=25)
    ub, tb, beta0, x, A = sph.encode(sp, oc, mults, 21, c.encode)
    assert A.shape == (n, L * M) and ub.size == 4 * c.K
    rng = np.random.RandomState(21)
    y = x + np.sqrt(sp["P"] / 6.0) * rng.randn(n)
    trace = []
    ref_beta, ref_s = sph.sparc_amp(y, sp, dp, A, trace)
    beta, s_, tau = sparc_new.sparc_amp(y, sp, dp, A, precision="tc", return_trace=True)
    ref_tau = np.array([t["tau_sqr"] for t in trace])
    assert np.max(np.abs(tau - ref_tau) / ref_tau) < 1e-4
    assert np.array_equal(s_.reshape(L, M).argmax(1), ref_s.reshape(L, M).argmax(1))
    ref_sep = sph.sparc_ldpc_decode(y, sp, oc, dp, True, lengths, A)
    ref_int = sph.integrated_decoder(y, sp, oc, dp, A)
    assert np.array_equal(sparc_new.sparc_ldpc_decode(y, sp, lp, dp, True, lengths, A, precision="tc"), ref_sep)
    assert np.array_equal(sparc_new.integrated_decoder(y, sp, lp, dp, A, precision="tc"), ref_int)
    assert np.array_equal(ref_int, ub)                              # SNR 6: the top of the reference's sweep decodes cleanly


def test_config3_full_size_tensor_core_amp_vs_oracle():
    """BASELINE config 3 at full size (L = 512, M = 64, n = 3072: A 3072 x 32768, P = 19.44,
    t_max = 25): the tcgen05 AMP path against the numpy oracle on the same A and y --
    per-iteration tau^2 and the final beta within 1e-4, identical section decisions."""
    L, M, n, P = 512, 64, 3072, 19.44
    rng = np.random.default_rng(0)
    A = rng.normal(0, 1 / np.sqrt(n), (n, L * M))
    sq = np.sqrt(n * P / L)
    sp, dp = dict(P=P, R=1.0, L=L, M=M), dict(t_max=25)
    ys, idx0 = [], []
    for f in range(3):
        idx = rng.integers(0, M, L)
        beta0 = np.zeros(L * M)
        beta0[np.arange(L) * M + idx] = sq
        ys.append(A @ beta0 + np.sqrt(P / 6.0) * rng.standard_normal(n))
        idx0.append(idx)
    ys = np.stack(ys)
    beta, s, tau = sparc_new.sparc_amp(ys, sp, dp, A, precision="tc", return_trace=True)
    for f in range(3):
        trace = []
        ref_beta, ref_s = sph.sparc_amp(ys[f], sp, dp, A, trace)
        ref_tau = np.array([t["tau_sqr"] for t in trace])
        assert np.max(np.abs(tau[f] - ref_tau) / ref_tau) < 1e-4
        assert np.max(np.abs(beta[f] - ref_beta)) < 1e-4 * sq
        dec, ref_dec = s[f].reshape(L, M).argmax(1), ref_s.reshape(L, M).argmax(1)
        assert np.array_equal(dec, ref_dec) and np.mean(dec != idx0[f]) < 0.01


def test_two_stream_drive_equals_one_stream(gold, monkeypatch):
    """Large tensor-core batches run as two half-batches on two CUDA streams (sparc_new._drive).
    Frames are independent and every kernel is deterministic, so the result must be bit-identical
    to decoding the same two half-batches one after the other on one stream.  (Against the
    undivided batch only the K-split of the GEMMs, i.e. the fp32 summation order, differs: tau^2
    agrees to 1e-5 and the frame error count is the same up to frames at the decoding threshold.)"""
    g, meta = gold
    m = meta["z54_m2_hi"]
    pre, c, sp, dp, y, A = setup_case(g, m, "z54_m2_hi", 0)
    lp = dict(standard=m["standard"], rate=m["rate"], z=m["z"])
    lengths = dict(k_ldpc=c.K * m["mults"], mults=m["mults"], L_unprotected=0)
    rng = np.random.RandomState(2)
    F = 300
    ys = np.stack([g[pre + "x"] + np.sqrt(sp["P"] / 5.0) * rng.randn(len(y)) for _ in range(F)])
    cut = (F // 2 + 127) // 128 * 128                      # as sparc_new._drive cuts
    monkeypatch.setattr(sparc_new, "FUSED_TC_AMP", False)  # the plain AMP loop through _drive too (fused: one C call)

    def all_decoders(batch):
        return dict(
            amp=sparc_new.sparc_amp(batch, sp, dp, A, precision="tc", return_trace=True),
            sep=sparc_new.sparc_ldpc_decode(batch, sp, lp, dp, True, lengths, A, precision="tc"),
            naive=sparc_new.naively_integrated_decoder(batch, sp, lp, dp, A, precision="tc"),
            integ=sparc_new.integrated_decoder(batch, sp, lp, dp, A, precision="tc", return_trace=True))

    monkeypatch.setattr(sparc_new, "TWO_STREAM_MIN_FRAMES", 64)
    two = all_decoders(ys)
    monkeypatch.setattr(sparc_new, "TWO_STREAM_MIN_FRAMES", 10 ** 9)
    a, b, whole = all_decoders(ys[:cut]), all_decoders(ys[cut:]), all_decoders(ys)
    for k in ("sep", "naive"):
        assert np.array_equal(two[k], np.concatenate([a[k], b[k]]))
    for k, parts in (("amp", 3), ("integ", 3)):
        for i in range(parts):
            assert np.array_equal(two[k][i], np.concatenate([a[k][i], b[k][i]])), (k, i)
    truth = g[pre + "user_bits"]
    assert np.allclose(two["amp"][2], whole["amp"][2], rtol=1e-5)
    for k in ("sep", "naive"):
        # 1/3 of these frames sit at the decoding threshold: a different fp32 summation order may flip a few of them
        assert abs(int((two[k] != truth).any(axis=1).sum()) - int((whole[k] != truth).any(axis=1).sum())) <= 3


@pytest.mark.parametrize("L,M,n,F", [(64, 64, 384, 5), (48, 2, 96, 130), (24, 128, 256, 3), (40, 16, 320, 257),
                                     (256, 4, 512, 64)])
def test_fused_tensor_core_amp_equals_kernel_chain(monkeypatch, L, M, n, F):
    """lsb_tc_amp_dev_f32 (A beta -> residual -> A^T z with the denoiser as the GEMM epilogue, the fp16 operand
    splits written by the producers) against the unfused chain of the same products and the separate residual /
    denoiser kernels, and against the numpy oracle: section sizes 2 .. 128, frame counts that do not fill a
    128-frame tile, beta lengths that do not fill a 128-row tile."""
    P = 12.0
    rng = np.random.default_rng(L + M)
    A = rng.normal(0, 1 / np.sqrt(n), (n, L * M))
    sq = np.sqrt(n * P / L)
    sp, dp = dict(P=P, R=1.0, L=L, M=M), dict(t_max=12)
    beta0 = np.zeros((F, L * M))
    idx = rng.integers(0, M, (F, L))
    beta0[np.arange(F)[:, None], np.arange(L)[None, :] * M + idx] = sq
    ys = beta0 @ A.T + np.sqrt(P / 8.0) * rng.standard_normal((F, n))
    monkeypatch.setattr(sparc_new, "FUSED_TC_AMP", True)
    before = _lib.lib().lsb_launch_count()
    beta, s, tau = sparc_new.sparc_amp(ys, sp, dp, A, precision="tc", return_trace=True)
    assert _lib.lib().lsb_launch_count() - before <= 3 * dp["t_max"] + 8          # three launches per iteration
    monkeypatch.setattr(sparc_new, "FUSED_TC_AMP", False)
    beta_u, s_u, tau_u = sparc_new.sparc_amp(ys, sp, dp, A, precision="tc", return_trace=True)
    assert tau.shape == tau_u.shape == (F, dp["t_max"])
    assert np.max(np.abs(tau - tau_u) / tau_u) < 2e-5
    # beta of a section at its decision threshold amplifies the last-bit differences of s by sqrt(nP_l) / tau^2
    # (the small codes of this test do not converge cleanly: both chains sit 3e-4 from the float64 oracle there)
    assert np.max(np.abs(beta - beta_u)) < 5e-4 * sq and np.max(np.abs(s - s_u)) < 1e-4 * sq
    assert np.array_equal(s.reshape(F, L, M).argmax(2), s_u.reshape(F, L, M).argmax(2))
    for f in (0, F - 1):
        trace = []
        ref_beta, ref_s = sph.sparc_amp(ys[f], sp, dp, A, trace)
        ref_tau = np.array([t["tau_sqr"] for t in trace])
        assert np.max(np.abs(tau[f] - ref_tau) / ref_tau) < 1e-4
        assert np.max(np.abs(beta[f] - ref_beta)) < 5e-4 * sq          # 1e-4 at full size: test_config3_full_size_...
        assert np.array_equal(s[f].reshape(L, M).argmax(1), ref_s.reshape(L, M).argmax(1))


def test_dense_amp_host_entry_equals_device_path():
    """lsb_tc_amp_host_f32 (host y in, MAP section decisions out) against argmax of the device path's s."""
    L, M, n, F, P = 64, 32, 320, 70, 12.0
    rng = np.random.default_rng(11)
    A = rng.normal(0, 1 / np.sqrt(n), (n, L * M))
    sq = np.sqrt(n * P / L)
    sp, dp = dict(P=P, R=1.0, L=L, M=M), dict(t_max=15)
    idx0 = rng.integers(0, M, (F, L))
    beta0 = np.zeros((F, L * M))
    beta0[np.arange(F)[:, None], np.arange(L)[None, :] * M + idx0] = sq
    ys = (beta0 @ A.T + np.sqrt(P / 8.0) * rng.standard_normal((F, n))).astype(np.float32)
    idx_h = sparc_new.sparc_amp_decide_host(ys, sp, dp, A)
    beta, s = sparc_new.sparc_amp(ys.astype(np.float64), sp, dp, A, precision="tc")
    assert np.array_equal(idx_h, s.reshape(F, L, M).argmax(2))
    assert (idx_h != idx0).mean() < 0.05

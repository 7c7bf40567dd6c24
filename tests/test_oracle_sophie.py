"""CPU: the sparc_sophie oracle against the reference's golden vectors (and, in the
build container, bit for bit against the imported reference)."""
import json
import os

import numpy as np
import pytest

import sophie_oracle as sph
from conftest import GOLDEN
from ldpc_sparc_b200 import ldpc, sparc_new


@pytest.fixture(scope="module")
def gold():
    g = np.load(os.path.join(GOLDEN, "sophie_reference.npz"))
    return g, json.loads(str(g["meta_json"]))


def setup_case(g, m, name, si):
    pre = "%s_s%d_" % (name, si)
    c = ldpc.code(m["standard"], m["rate"], m["z"])
    M = 2 ** m["logM"]
    L = c.N * m["mults"] // m["logM"]
    sp = dict(P=m["P"], R=m["R"], L=L, M=M)
    y = g[pre + "y"]
    A = sparc_new.create_design_matrix(L, M, len(y), m["seeds"][si])
    chk = np.array([A.sum(), np.abs(A).sum(), A[0, 0], A[-1, -1], A[A.shape[0] // 2, 7]])
    if not np.array_equal(chk, g[pre + "A_checksum"]):
        pytest.skip("numpy default_rng stream differs from the one the goldens were made with")
    return pre, c, sp, dict(t_max=m["t_max"]), y, A


def test_oracle_matches_reference_golden(gold):
    g, meta = gold
    for name, m in meta.items():
        for si in range(len(m["seeds"])):
            pre, c, sp, dp, y, A = setup_case(g, m, name, si)
            oc = sph.Code(c)
            lengths = dict(k_ldpc=c.K * m["mults"], mults=m["mults"], L_unprotected=0)
            trace = []
            beta, s = sph.sparc_amp(y, sp, dp, A, trace)
            assert np.array_equal(beta, g[pre + "amp_beta"]) and np.array_equal(s, g[pre + "amp_s"])
            assert np.array_equal([t["tau_sqr"] for t in trace], g[pre + "amp_tau_sqr"])
            assert np.array_equal(sph.sparc_ldpc_decode(y, sp, oc, dp, True, lengths, A), g[pre + "bits_separate"])
            assert np.array_equal(sph.naively_integrated_decoder(y, sp, oc, dp, A), g[pre + "bits_naive"])
            tr = []
            assert np.array_equal(sph.integrated_decoder(y, sp, oc, dp, A, tr), g[pre + "bits_integrated"])
            assert np.array_equal([t["onsager_sum"] for t in tr][1:], g[pre + "int_onsager_sum"][:len(tr) - 1])


def test_glue_known_answers():
    """The reference's own KATs: test_S_k_mapping.py:31-39, test_beta_estimate_to_bp_probs.py:25-52."""
    assert sparc_new.S_k_mapping(4) == [[0, 1], [0, 2]] == sph.S_k_mapping(4)
    assert sparc_new.S_k_mapping(8) == [[0, 1, 2, 3], [0, 1, 4, 5], [0, 2, 4, 6]] == sph.S_k_mapping(8)
    assert sph.S_k_mapping(16)[3] == [0, 2, 4, 6, 8, 10, 12, 14]
    beta = np.array([1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0], dtype=float)     # sections -> indices 0, 2, 1
    assert sph.beta_estimate_to_bp_probs(beta, 3, 4, 1.0).tolist() == [1, 1, 0, 1, 1, 0]
    p = np.array([0.9, 0.2, 0.5, 0.5])
    b = sph.bp_output_to_beta_estimate(p, 2, 4, 2.0)
    assert np.allclose(b[:4], 2.0 * np.array([0.9 * 0.2, 0.9 * 0.8, 0.1 * 0.2, 0.1 * 0.8]))


@pytest.mark.reference
def test_oracle_bit_exact_vs_imported_reference():
    import ref_shims
    R = ref_shims.load()
    RN, RSN = R["sparc_new"], R["sparc_sim_new"]
    c_ref = R["ldpc"].code("802.16", "1/2", 27)
    oc = sph.Code(c_ref)
    sp = dict(P=19.44, R=1.0, L=324, M=4)
    lp = dict(standard="802.16", rate="1/2", z=27)
    lengths = dict(k_ldpc=c_ref.K, mults=1, L_unprotected=0)
    dp = dict(t_max=6)
    ub, tb, beta0, x, A = RN.sparc_ldpc_encode(sp, lp, lengths, True, 9)
    y = RSN.awgn_channel(x, 19.44 / 3.4, 9)
    assert np.array_equal(RN.sparc_ldpc_decode(y, sp, lp, dp, True, lengths, A), sph.sparc_ldpc_decode(y, sp, oc, dp, True, lengths, A))
    assert np.array_equal(RN.naively_integrated_decoder(y, sp, lp, dp, A), sph.naively_integrated_decoder(y, sp, oc, dp, A))
    assert np.array_equal(RN.integrated_decoder(y, sp, lp, dp, A), sph.integrated_decoder(y, sp, oc, dp, A))

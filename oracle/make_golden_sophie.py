"""sparc_sophie golden vectors from the UNMODIFIED reference (see make_golden.py).

For each case: sparc_ldpc_encode -> awgn_channel -> {sparc_amp, sparc_ldpc_decode,
naively_integrated_decoder, integrated_decoder} under oracle/ref_shims.py, with
sparc_amp_single_it / eta / differentiated_eta_calc wrapped (not edited) to record
the per-iteration tau^2, post-BP beta and Onsager sums.  The design matrix A is NOT
stored (megabytes): it is regenerated from the seed by create_design_matrix and
checked against the stored checksum before use.
"""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (standard, rate, z, logM, mults, R, snr, t_max)
    "z27_m2": ("802.16", "1/2", 27, 1, 1, 1.0, 3.0, 8),       # M=2 as performance_plots_general.py:27-46, small z
    "z27_m4": ("802.16", "1/2", 27, 2, 1, 1.0, 3.2, 8),
    "n11_m8": ("802.11n", "1/2", 27, 3, 2, 1.0, 3.5, 6),
    "z54_m2_hi": ("802.16", "1/2", 54, 1, 1, 1.0, 5.0, 10),
}
SEEDS = (1, 2)


def run_case(ref, std, rate, z, logM, mults, R, snr, t_max, seed):
    RN, RSN = ref["sparc_new"], ref["sparc_sim_new"]
    c = ref["ldpc"].code(std, rate, z)
    M = 2 ** logM
    L = c.N * mults // logM
    sp = dict(P=19.44, R=R, L=L, M=M)
    lp = dict(standard=std, rate=rate, z=z)
    lengths = dict(k_ldpc=c.K * mults, mults=mults, L_unprotected=0)
    dp = dict(t_max=t_max)
    awgn_var = sp["P"] / snr
    user_bits, total_bits, beta0, x, A = RN.sparc_ldpc_encode(sp, lp, lengths, True, seed)
    y = RSN.awgn_channel(x, awgn_var, seed)
    out = dict(user_bits=user_bits.astype(np.uint8), total_bits=total_bits.astype(np.uint8), x=x, y=y,
               A_checksum=np.array([A.sum(), np.abs(A).sum(), A[0, 0], A[-1, -1], A[A.shape[0] // 2, 7]]))
    # plain AMP with per-iteration tau^2
    taus = []
    orig_mmse = RN.msg_vector_mmse_estimator

    def spy_mmse(s, tau_sqr, n, P_l, M_):
        taus.append(float(tau_sqr))
        return orig_mmse(s, tau_sqr, n, P_l, M_)

    RN.msg_vector_mmse_estimator = spy_mmse
    try:
        beta_soft, s = RN.sparc_amp(y, sp, dp, A)
    finally:
        RN.msg_vector_mmse_estimator = orig_mmse
    out.update(amp_tau_sqr=np.array(taus), amp_beta=beta_soft, amp_s=s)
    out["bits_separate"] = np.asarray(RN.sparc_ldpc_decode(y, sp, lp, dp, True, lengths, A)).astype(np.uint8)
    # naive decoder: tau^2 per iteration
    taus = []
    orig_it = RN.sparc_amp_single_it

    def spy_it(*a):
        r = orig_it(*a)
        taus.append(float(r[2]))
        return r

    RN.sparc_amp_single_it = spy_it
    try:
        out["bits_naive"] = np.asarray(RN.naively_integrated_decoder(y, sp, lp, dp, A)).astype(np.uint8)
    finally:
        RN.sparc_amp_single_it = orig_it
    out["naive_tau_sqr"] = np.array(taus)
    # integrated decoder: tau^2 and Onsager sum per iteration
    taus, sums = [], []
    orig_eta, orig_d = RN.eta, RN.differentiated_eta_calc

    def spy_eta(s_, tau_sqr, *a):
        taus.append(float(tau_sqr))
        return orig_eta(s_, tau_sqr, *a)

    def spy_d(*a):
        r = orig_d(*a)
        sums.append(float(np.sum(r)))
        return r

    RN.eta, RN.differentiated_eta_calc = spy_eta, spy_d
    try:
        out["bits_integrated"] = np.asarray(RN.integrated_decoder(y, sp, lp, dp, A)).astype(np.uint8)
    finally:
        RN.eta, RN.differentiated_eta_calc = orig_eta, orig_d
    out["int_tau_sqr"] = np.array(taus)
    out["int_onsager_sum"] = np.array(sums)
    return out


def make_sophie(ref):
    out, meta = {}, {}
    for name, (std, rate, z, logM, mults, R, snr, t_max) in CASES.items():
        meta[name] = dict(standard=std, rate=rate, z=z, logM=logM, mults=mults, R=R, snr=snr, t_max=t_max, P=19.44,
                          seeds=list(SEEDS))
        for si, seed in enumerate(SEEDS):
            r = run_case(ref, std, rate, z, logM, mults, R, snr, t_max, seed)
            for k, v in r.items():
                out["%s_s%d_%s" % (name, si, k)] = v
            ub = r["user_bits"]
            print(name, seed, "ber separate/naive/integrated: %.4f %.4f %.4f" % tuple(
                float(np.mean(r[k] != ub)) for k in ("bits_separate", "bits_naive", "bits_integrated")))
    out["meta_json"] = np.array(json.dumps(meta))
    path = os.path.join(GOLD, "sophie_reference.npz")
    np.savez_compressed(path, **out)
    print("sophie_reference.npz: %.1f kB" % (os.path.getsize(path) / 1e3))

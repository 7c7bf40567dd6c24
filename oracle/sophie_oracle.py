"""sophie_oracle.py -- CPU ORACLE (test infrastructure, NOT the product).

numpy restatement of the reference's concatenated LDPC-SPARC path
(sparc_sophie/sparc_new.py): dense Gaussian design, AMP, section<->bit glue and the
three AMP+BP decoders.  Loops that are pure-Python in the reference are vectorised
over sections but keep the reference's order of floating-point operations inside a
section, so results are bit-identical (checked against the imported reference by
tests/test_oracle_sophie.py and pinned by tests/golden/sophie_reference.npz, both
produced under oracle/ref_shims.py).  BP runs through oracle/ldpc_oracle.c.
Only tests/, smoke() and bench.py's CPU legs may import this file.
"""
import numpy as np

import ldpc_oracle as orc


class Code:
    """Graph + sizes of an LDPC code for the oracle (built from any object with
    vdeg/cdeg/intrlv/N/K, e.g. the façade's or the reference's ``code``)."""

    def __init__(self, c):
        self.graph = orc.Graph(c.vdeg, c.cdeg, c.intrlv)
        self.N, self.K = int(c.N), int(c.K)

    def decode(self, ch, max_it):
        return orc.decode(self.graph, ch, max_it)


def create_design_matrix(L, M, n, rand_seed):
    """sparc_new.py:1284-1294."""
    rng = np.random.default_rng(rand_seed)
    return rng.normal(loc=0, scale=1 / np.sqrt(n), size=(n, M * L))


def bin_arr_2_msg_vector(bits, M, n, P_l):
    """sparc_new.py:1298-1318."""
    logM = int(np.log2(M))
    L = bits.size // logM
    idx = bits.reshape(L, logM).astype(np.int64).dot(1 << np.arange(logM)[::-1])
    beta = np.zeros(L * M)
    beta[np.arange(L) * M + idx] = np.sqrt(n * P_l)
    return beta


def msg_vector_2_bin_arr(beta, M):
    """sparc_new.py:1320-1341 (int_2_bin_arr with its numpy-1.x meaning)."""
    logM = int(round(np.log2(M)))
    L = beta.size // M
    rows, idx = np.nonzero(beta.reshape(L, M))
    assert np.array_equal(rows, np.arange(L))
    return ((idx[:, None] >> np.arange(logM)[::-1]) & 1).astype(bool).ravel()


def mmse_estimator(s, tau_sqr, n, P_l, M):
    """sparc_new.py:1040-1066 (float64, global max shift)."""
    x = np.sqrt(n * P_l) * (s / tau_sqr)
    top = np.exp(x - x.max(), dtype=np.float64)
    bot = top.reshape(-1, M).sum(axis=1).repeat(M)
    return (np.sqrt(n * P_l) * (top / bot)).astype(np.float64)


def map_estimator(s, M, sqrt_nP_l):
    """sparc_new.py:1099-1116."""
    L = s.size // M
    beta = np.zeros((L, M))
    beta[np.arange(L), s.reshape(L, M).argmax(axis=1)] = sqrt_nP_l
    return beta.ravel()


def S_k_mapping(M):
    """sparc_new.py:1140-1160: S_k[b] = indices whose bit b (b = 0 is the MSB) is 0."""
    logM = int(np.log2(M))
    S = [[] for _ in range(logM)]
    for i in range(logM):
        b = logM - 1 - i
        k = 0
        while k < M:
            S[b].extend(range(k, k + 2 ** i))
            k += 2 ** (i + 1)
    return S


def beta_estimate_to_bp_probs(beta, L, M, sqrt_nP_l):
    """sparc_new.py:1118-1138: P(bit b = 0) = sum_{j in S_b} beta_j / sqrt(nP_l), summed in
    increasing j (the reference's loop order)."""
    logM = int(np.log2(M))
    bs = beta.reshape(L, M)
    probs = np.zeros((L, logM))
    for b, idxs in enumerate(S_k_mapping(M)):
        for j in idxs:
            probs[:, b] += bs[:, j] / sqrt_nP_l
    return probs.reshape(L * logM)


def bp_output_to_beta_estimate(probs, L, M, sqrt_nP_l):
    """sparc_new.py:1260-1279: beta_i = sqrt(nP_l) prod_b (p_b if bit b of i is 0 else 1-p_b)."""
    logM = int(np.log2(M))
    p = probs.reshape(L, logM)
    amp = np.ones((L, M))
    for i in range(M):
        for j in range(logM):
            bit = (i >> (logM - 1 - j)) & 1
            amp[:, i] = amp[:, i] * p[:, j] if bit == 0 else amp[:, i] * (1 - p[:, j])
    return amp.reshape(L * M) * sqrt_nP_l


def ldpc_bp(probs, c, num_its, hard):
    """sparc_new.py:1162-1193.  c: Code.  Returns (probs_out, hard_bits)."""
    eps = 1e-15
    p = np.clip(probs, eps, 1 - eps)
    LLR = np.log(p) - np.log(1 - p)
    assert len(LLR) % c.N == 0
    blocks = LLR.reshape(-1, c.N)
    app = np.stack([c.decode(ch, num_its)[0] for ch in blocks])
    if hard:
        return 0, (app[:, :c.K].ravel() < 0).astype(int)
    with np.errstate(over="ignore", invalid="ignore"):
        out = np.exp(app.ravel()) / (1 + np.exp(app.ravel()))
    return out, 0


def sparc_amp(y, sparc_params, decode_params, A, trace=None):
    """sparc_new.py:885-912.  Returns (beta_soft, s)."""
    P, L, M = sparc_params["P"], sparc_params["L"], sparc_params["M"]
    n = len(y)
    P_l = P / L
    AT = A.T
    beta = np.zeros(L * M)
    z = y
    for t in range(decode_params["t_max"]):
        if t > 0:
            Ab = np.dot(A, beta)
            onsager = (z / tau_sqr) * (P - (np.sum(beta ** 2)) / n)
            z = y - Ab + onsager
        s = beta + np.dot(AT, z)
        tau_sqr = np.sum(z ** 2) / n
        beta = mmse_estimator(s, tau_sqr, n, P_l, M)
        if trace is not None:
            trace.append(dict(tau_sqr=tau_sqr, beta=beta.copy(), s=s.copy()))
    return beta, s


def sparc_amp_single_it(sparc_params, y, A, AT, beta, z, tau_sqr):
    """sparc_new.py:975-990."""
    P, L, M = sparc_params["P"], sparc_params["L"], sparc_params["M"]
    P_l = P / L
    n = len(y)
    onsager = (z / tau_sqr) * (P - (np.sum(beta ** 2)) / n)
    z = y - np.dot(A, beta) + onsager
    s = beta + np.dot(AT, z)
    tau_sqr = np.sum(z ** 2) / n
    return mmse_estimator(s, tau_sqr, n, P_l, M), z, tau_sqr


def sparc_ldpc_decode(y, sparc_params, c, decode_params, ldpc_bool, lengths, A):
    """sparc_new.py:53-82 (L_unprotected = 0 is the only working case of the reference)."""
    P, L, M = sparc_params["P"], sparc_params["L"], sparc_params["M"]
    n = len(y)
    sq = np.sqrt(n * (P / L))
    beta, s = sparc_amp(y, sparc_params, decode_params, A)
    if ldpc_bool:
        Lu = lengths["L_unprotected"]
        unprot = map_estimator(s, M, sq)[:int(Lu * M)]
        unprot_bits = msg_vector_2_bin_arr(unprot, M)
        probs = beta_estimate_to_bp_probs(beta[int(Lu * M):], L, M, sq)
        _, prot_bits = ldpc_bp(probs, c, 200, True)
        return np.concatenate((unprot_bits, prot_bits))
    return msg_vector_2_bin_arr(map_estimator(s, M, sq), M)


def naively_integrated_decoder(y, sparc_params, c, decode_params, A, trace=None):
    """sparc_new.py:257-282."""
    P, L, M = sparc_params["P"], sparc_params["L"], sparc_params["M"]
    n = len(y)
    sq = np.sqrt(n * (P / L))
    t_max = decode_params["t_max"]
    beta, z, tau_sqr = np.zeros(L * M), 0, 1
    AT = A.T
    for i in range(t_max):
        beta, z, tau_sqr = sparc_amp_single_it(sparc_params, y, A, AT, beta, z, tau_sqr)
        probs = beta_estimate_to_bp_probs(beta, L, M, sq)
        if i != t_max - 1:
            probs, _ = ldpc_bp(probs, c, 6, False)
            beta = bp_output_to_beta_estimate(probs, L, M, sq)
            if trace is not None:
                trace.append(dict(tau_sqr=tau_sqr, beta=beta.copy()))
        else:
            _, bits = ldpc_bp(probs, c, 200, True)
    return bits


def differentiated_eta(beta, vk, vk_0, alpha, tau_sqr, L, M, n, P_l):
    """sparc_new.py:824-841 with sub_term :871-883, vectorised over sections, same
    operation order inside a section."""
    logM = int(np.log2(M))
    S = S_k_mapping(M)
    v = vk.reshape(L, logM)
    a = alpha.reshape(L, M)
    v0 = np.clip(vk_0.reshape(L, logM), 1e-10, 1 - 1e-10)
    c = np.sqrt(n * P_l) / tau_sqr
    main = np.zeros((L, M))
    for i in range(M):
        for k in range(logM):
            sum_term = np.zeros(L)
            for q in S[k]:
                if q == i:
                    sum_term += a[:, q] * c * (1 - a[:, q])
                else:
                    sum_term += a[:, q] * c * (-a[:, i])
            sub = (1 / (v0[:, k] * (1 - v0[:, k]))) * sum_term
            if (i >> (logM - 1 - k)) & 1:
                main[:, i] -= v[:, k] * sub
            else:
                main[:, i] += (1 - v[:, k]) * sub
    return beta * main.reshape(L * M)


def eta(s, tau_sqr, n, P_l, M, L, c, num_its, num_its_final, hard):
    """sparc_new.py:709-735."""
    sq = np.sqrt(n * P_l)
    weighted = mmse_estimator(s, tau_sqr, n, P_l, M)
    alpha = weighted / sq
    vk_0 = beta_estimate_to_bp_probs(weighted, L, M, sq)
    if hard:
        vk, bits = ldpc_bp(vk_0, c, num_its_final, True)
        beta = np.zeros(L * M)
    else:
        vk, bits = ldpc_bp(vk_0, c, num_its, False)
        beta = bp_output_to_beta_estimate(vk, L, M, sq)
    return alpha, vk_0, vk, beta, bits


def integrated_decoder(y, sparc_params, c, decode_params, A, trace=None):
    """sparc_new.py:472-502."""
    P, L, M = sparc_params["P"], sparc_params["L"], sparc_params["M"]
    n = len(y)
    P_l = P / L
    t_max = decode_params["t_max"]
    beta, z, d_eta = np.zeros(L * M), 0, np.zeros(n)
    AT = A.T
    for t in range(t_max):
        if t != 0:
            d_eta = differentiated_eta(beta, vk, vk_0, alpha, tau_sqr, L, M, n, P_l)
        z = y - np.dot(A, beta) + (z / n) * (np.sum(d_eta))
        s = np.dot(AT, z) + beta
        tau_sqr = np.sum(z ** 2) / n
        if t != t_max - 1:
            alpha, vk_0, vk, beta, _ = eta(s, tau_sqr, n, P_l, M, L, c, 6, 200, False)
            if trace is not None:
                trace.append(dict(tau_sqr=tau_sqr, beta=beta.copy(), onsager_sum=float(np.sum(d_eta))))
        else:
            _, _, _, _, bits = eta(s, tau_sqr, n, P_l, M, L, c, 6, 200, True)
    return bits


def encode(sparc_params, c, mults, rng_seed, enc):
    """sparc_ldpc_encode (sparc_new.py:15-51) with L_unprotected = 0.  `enc(info)->codeword`
    is the LDPC encoder (host logic, e.g. the façade's code.encode)."""
    P, R, L, M = sparc_params["P"], sparc_params["R"], sparc_params["L"], sparc_params["M"]
    logM = int(np.log2(M))
    rng = np.random.default_rng(rng_seed)
    user_bits = rng.integers(0, 2, size=int(c.K * mults))
    total = np.concatenate([enc(b) for b in np.array_split(user_bits, mults)]).astype(bool)
    assert total.size == L * logM
    n = int(total.size / R)
    beta0 = bin_arr_2_msg_vector(total, M, n, P / L)
    A = create_design_matrix(L, M, n, rng_seed)
    return user_bits, total, beta0, np.dot(A, beta0), A

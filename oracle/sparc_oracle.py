"""sparc_oracle.py -- CPU ORACLE (test infrastructure, NOT the product).

numpy restatement of the reference's real-valued, unmodulated SPARC path
(sparc_public/sparc.py), each function citing the lines it follows.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.

Parity pinning: tests/test_oracle_sparc.py checks it (a) against the committed
golden vectors tests/golden/sparc_*.npz produced by the unmodified reference
(oracle/make_golden_sparc.py), (b) in the build container, bit for bit against the
imported reference on fresh seeds, and (c) checks the implicit DCT operators
against the explicit closed-form matrix.  Third-party arithmetic: the reference
calls scipy.fftpack.dct/idct (sparc.py:6,691,698; SciPy unpinned in
requirements.txt); this file calls the same routines through scipy.fft, whose
result at that boundary is pinned by (c).
"""
import numpy as np
from scipy.fft import dct, idct


# ------------------------------------------------------------ base matrix W
def pa_iterative(P, sigma2, B, R_PA):
    """sparc.py:516-533."""
    Q = np.zeros(B)
    for b in range(B):
        phi = sigma2 + P - Q.mean()
        p_block = 2 * np.log(2) * R_PA * phi
        p_spread = (B * P - Q.sum()) / (B - b)
        if p_block > p_spread:
            Q[b:b + 1] = p_block
        else:
            Q[b:] = p_spread
            break
    Q /= Q.mean() / P
    return Q


def sc_basic(Q, omega, Lambda):
    """sparc.py:535-568."""
    Lr = Lambda + omega - 1
    if Q.ndim == 0:
        W = np.zeros((Lr, Lambda))
        for c in range(Lambda):
            W[c:c + omega, c] = Q * Lr / omega
    else:
        B = Q.size
        W = np.zeros((Lr, Lambda * B))
        for c in range(Lambda):
            for r in range(c, c + omega):
                W[r, c * B:(c + 1) * B] = Q * Lr / omega
    assert np.isclose(W.mean(), np.mean(Q))
    return W


def create_base_matrix(P, power_allocated=False, spatially_coupled=False, **kw):
    """sparc.py:570-589."""
    Q = np.array(P) if not power_allocated else pa_iterative(P, kw["awgn_var"], kw["B"], kw["R"] * kw["R_PA_ratio"])
    return Q if not spatially_coupled else sc_basic(Q, kw["omega"], kw["Lambda"])


def code_dims(code_params, W):
    """n, R_actual as sparc_encode fixes them (sparc.py:27,39-45)."""
    L, M, R = code_params["L"], code_params["M"], code_params["R"]
    K = code_params["K"] if code_params.get("modulated") else 1
    bit_len = int(round(L * np.log2(K * M)))
    n = int(round(bit_len / R))
    if W.ndim == 2:
        Lr = W.shape[0]
        n = int(round(n / Lr)) * Lr
    return n, bit_len / n


# ------------------------------------------------------- implicit design matrix
class Design:
    """Sub-sampled DCT design (sparc.py:703-880, real case).

    Block (r, c) of A is sqrt(W[r,c]/L) * sqrt(2) * cos(pi*o0[i]*(2*o1[j]+1)/(2w)),
    with o0/o1 the first Mr/Mc entries of two arange(1, w) arrays that are
    shuffled IN PLACE, cumulatively, once per non-zero block, by one
    RandomState(rand_seed) (generate_ordering, sparc.py:735-775)."""

    def __init__(self, W, L, M, n, rand_seed, csparc=False):
        W = np.asarray(W, dtype=float)
        self.W, self.L, self.M, self.n, self.csparc = W, L, M, n, csparc
        if W.ndim == 0:
            self.Lr, self.Lc = 1, 1
            Wm = W.reshape(1, 1)
        elif W.ndim == 1:
            self.Lr, self.Lc = 1, W.size
            Wm = W.reshape(1, -1)
        else:
            self.Lr, self.Lc = W.shape
            Wm = W
        assert (L * M) % self.Lc == 0 and n % self.Lr == 0
        self.Wm = Wm
        self.Mr, self.Mc = n // self.Lr, L * M // self.Lc
        if not csparc:
            self.w = 2 ** int(np.ceil(np.log2(max(self.Mr + 1, self.Mc + 1))))
            idx0 = np.arange(1, self.w, dtype=np.uint32)
            idx1 = np.arange(1, self.w, dtype=np.uint32)
        else:      # sub-sampled DFT design: rows/columns 0 and w/2 are never used (sparc.py:747-750)
            self.w = 2 ** int(np.ceil(np.log2(max(self.Mr + 2, self.Mc + 2))))
            idx0 = np.delete(np.arange(self.w, dtype=np.uint32), [0, self.w // 2])
            idx1 = np.delete(np.arange(self.w, dtype=np.uint32), [0, self.w // 2])
        rng = np.random.RandomState(rand_seed)
        self.order0 = np.zeros((self.Lr, self.Lc, self.Mr), dtype=np.uint32)
        self.order1 = np.zeros((self.Lr, self.Lc, self.Mc), dtype=np.uint32)
        for r in range(self.Lr):
            for c in range(self.Lc):
                if Wm[r, c] != 0:
                    rng.shuffle(idx0)
                    rng.shuffle(idx1)
                    self.order0[r, c] = idx0[:self.Mr]
                    self.order1[r, c] = idx1[:self.Mc]

    def Ab(self, x):
        """A @ x (sparc.py:786-789,811-819,851-862 with sub_dct.Ax :687-692)."""
        if self.csparc:
            return self._cab(np.asarray(x))
        x = np.asarray(x, dtype=float)
        assert x.size == self.Lc * self.Mc
        out = np.zeros(self.Lr * self.Mr)
        for r in range(self.Lr):
            for c in range(self.Lc):
                if self.Wm[r, c] != 0:
                    ext = np.zeros(self.w)
                    ext[self.order1[r, c]] = x[c * self.Mc:(c + 1) * self.Mc]
                    y = np.sqrt(self.w) * dct(ext, norm="ortho")
                    out[r * self.Mr:(r + 1) * self.Mr] += np.sqrt(self.Wm[r, c] / self.L) * y[self.order0[r, c]]
        return out

    def Az(self, y):
        """A.T @ y (sparc.py:791-794,821-829,864-875 with sub_dct.Ay :694-699)."""
        if self.csparc:
            return self._caz(np.asarray(y))
        y = np.asarray(y, dtype=float)
        assert y.size == self.Lr * self.Mr
        out = np.zeros(self.Lc * self.Mc)
        for r in range(self.Lr):
            for c in range(self.Lc):
                if self.Wm[r, c] != 0:
                    ext = np.zeros(self.w)
                    ext[self.order0[r, c]] = y[r * self.Mr:(r + 1) * self.Mr]
                    x = np.sqrt(self.w) * idct(ext, norm="ortho")
                    out[c * self.Mc:(c + 1) * self.Mc] += np.sqrt(self.Wm[r, c] / self.L) * x[self.order1[r, c]]
        return out

    def _cab(self, x):
        """Complex design, A @ x: zero-padded scatter, FFT, gather (sub_fft.Ax, sparc.py:632-637)."""
        out = np.zeros(self.Lr * self.Mr, dtype=complex)
        for r in range(self.Lr):
            for c in range(self.Lc):
                if self.Wm[r, c] != 0:
                    ext = np.zeros(self.w, dtype=complex)
                    ext[self.order1[r, c]] = x[c * self.Mc:(c + 1) * self.Mc]
                    out[r * self.Mr:(r + 1) * self.Mr] += np.sqrt(self.Wm[r, c] / self.L) * np.fft.fft(ext)[self.order0[r, c]]
        return out

    def _caz(self, y):
        """Complex design, A^H @ y (sub_fft.Ay, sparc.py:639-644)."""
        out = np.zeros(self.Lc * self.Mc, dtype=complex)
        for r in range(self.Lr):
            for c in range(self.Lc):
                if self.Wm[r, c] != 0:
                    ext = np.zeros(self.w, dtype=complex)
                    ext[self.order0[r, c]] = y[r * self.Mr:(r + 1) * self.Mr]
                    back = np.fft.fft(ext.conj()).conj()
                    out[c * self.Mc:(c + 1) * self.Mc] += np.sqrt(self.Wm[r, c] / self.L) * back[self.order1[r, c]]
        return out

    def dense(self):
        """Explicit A from the closed form (small sizes only)."""
        if self.csparc:
            A = np.zeros((self.Lr * self.Mr, self.Lc * self.Mc), dtype=complex)
            for r in range(self.Lr):
                for c in range(self.Lc):
                    if self.Wm[r, c] != 0:
                        o0 = self.order0[r, c].astype(float)[:, None]
                        o1 = self.order1[r, c].astype(float)[None, :]
                        A[r * self.Mr:(r + 1) * self.Mr, c * self.Mc:(c + 1) * self.Mc] = (
                            np.sqrt(self.Wm[r, c] / self.L) * np.exp(-2j * np.pi * o0 * o1 / self.w))
            return A
        A = np.zeros((self.Lr * self.Mr, self.Lc * self.Mc))
        for r in range(self.Lr):
            for c in range(self.Lc):
                if self.Wm[r, c] != 0:
                    o0 = self.order0[r, c].astype(float)[:, None]
                    o1 = self.order1[r, c].astype(float)[None, :]
                    A[r * self.Mr:(r + 1) * self.Mr, c * self.Mc:(c + 1) * self.Mc] = (
                        np.sqrt(self.Wm[r, c] / self.L) * np.sqrt(2) * np.cos(np.pi * o0 * (2 * o1 + 1) / (2 * self.w)))
        return A


# ------------------------------------------------------------------ PSK
def psk_constel(K):
    """sparc.py:225-239."""
    if K == 2:
        return np.array([1, -1])
    if K == 4:
        return np.array([1 + 0j, 0 + 1j, -1 + 0j, 0 - 1j])
    theta = 2 * np.pi * np.arange(K) / K
    return np.cos(theta) + 1j * np.sin(theta)


def gray_to_binary(g):
    """sparc.py:214-223."""
    mask = g >> 1
    while mask:
        g ^= mask
        mask >>= 1
    return g


# ------------------------------------------------------------------ estimators
def mmse_estimator(s, tau, M, K=1):
    """Section-wise MMSE estimator (sparc.py:402-465): K=1 (:429-432), K=2 (:433-441),
    K=4 (:442-453), K>4 (:454-459); complex observations halve tau first (:417-418);
    GLOBAL max shift, extended-precision exp, cast back (:462-465)."""
    if np.iscomplexobj(s):
        tau = tau / 2
    if K > 2:
        if K == 4:
            x, yv = s.real / tau, s.imag / tau
            mx = np.maximum(np.abs(x), np.abs(yv)).max()
            ex1, ex2 = np.exp(x - mx, dtype=np.longdouble), np.exp(-x - mx, dtype=np.longdouble)
            ey1, ey2 = np.exp(yv - mx, dtype=np.longdouble), np.exp(-yv - mx, dtype=np.longdouble)
            top = (ex1 - ex2) + 1j * (ey1 - ey2)
            bot = ((ex1 + ex2) + (ey1 + ey2)).reshape(-1, M).sum(axis=1).repeat(M)
        else:
            c = psk_constel(K)
            x = np.outer(s / tau, c.conj()).real
            tmp = np.exp(x - np.abs(x).max(), dtype=np.longdouble)
            top = np.dot(tmp, c)
            bot = tmp.sum(axis=1).reshape(-1, M).sum(axis=1).repeat(M)
        return (top / bot).astype(complex)
    x = s.real / tau
    if K == 1:
        top = np.exp(x - x.max(), dtype=np.longdouble)
        bot = top.reshape(-1, M).sum(axis=1).repeat(M)
    else:
        x_max = np.abs(x).max()
        e1 = np.exp(x - x_max, dtype=np.longdouble)
        e2 = np.exp(-x - x_max, dtype=np.longdouble)
        top = e1 - e2
        bot = (e1 + e2).reshape(-1, M).sum(axis=1).repeat(M)
    return (top / bot).astype(float)


def map_estimator(s, M, K=1):
    """sparc.py:485-487 (K=1), :488-492 (K=2: largest magnitude, its sign), :493-501 (K=4),
    :502-510 (K>4)."""
    L = s.size // M
    if K > 2:
        c = psk_constel(K)
        beta = np.zeros((L, M), dtype=complex)
        sm = s.reshape(L, M)
        if K == 4:
            idx = np.maximum(np.abs(sm.real), np.abs(sm.imag)).argmax(axis=1)
            k = np.rint(K * np.angle(sm[np.arange(L), idx]) / (2 * np.pi)).astype(int)
            k[k < 0] += K
            beta[np.arange(L), idx] = c[k]
        else:
            for l in range(L):
                corr = np.outer(sm[l].conj(), c).real
                j, k = np.unravel_index(corr.argmax(), corr.shape)
                beta[l, j] = c[k]
        return beta.ravel()
    beta = np.zeros((L, M))
    s2 = s.real.reshape(L, M)
    if K == 1:
        beta[np.arange(L), s2.argmax(axis=1)] = 1
    else:
        idx = np.abs(s2).argmax(axis=1)
        beta[np.arange(L), idx] = np.sign(s2[np.arange(L), idx])
    return beta.ravel()


# ------------------------------------------------------------------ bit <-> vector
def rnd_bin_arr(k, rand_seed):
    """sparc.py:174-180."""
    return np.random.RandomState(rand_seed).randint(2, size=k, dtype="bool")


def bin_arr_2_msg_vector(bits, M, K=1):
    """sparc.py:330-364: per section logM index bits (MSB first) and, for K=2, one value
    bit (2-PSK with Gray coding is the identity map: bit 0 -> +1, bit 1 -> -1, :227-253)."""
    logM = int(round(np.log2(M)))
    logK = int(round(np.log2(K)))
    sec = logM + logK
    L = bits.size // sec
    b = bits.reshape(L, sec)
    idx = b[:, :logM].astype(np.int64).dot(1 << np.arange(logM)[::-1])
    if K > 2:       # Gray-coded K-PSK value (psk_mod, sparc.py:241-269)
        word = b[:, logM:].astype(np.int64).dot(1 << np.arange(logK)[::-1])
        beta = np.zeros(L * M, dtype=complex)
        beta[np.arange(L) * M + idx] = psk_constel(K)[[gray_to_binary(int(v)) for v in word]]
        return beta
    beta = np.zeros(L * M)
    beta[np.arange(L) * M + idx] = 1 if K == 1 else 1.0 - 2.0 * b[:, logM]
    return beta


def msg_vector_2_bin_arr(beta, M, K=1):
    """sparc.py:366-400, with int_2_bin_arr's numpy-1.x meaning (:191-197)."""
    logM = int(round(np.log2(M)))
    L = beta.size // M
    rows, idx = np.nonzero(beta.reshape(L, M))
    assert np.array_equal(rows, np.arange(L))
    bits = ((idx[:, None] >> np.arange(logM)[::-1]) & 1).astype(bool)
    if K == 2:
        bits = np.concatenate([bits, (beta.reshape(L, M)[rows, idx] < 0)[:, None]], axis=1)
    elif K > 2:     # psk_demod (sparc.py:271-299): exact match against the constellation, Gray index
        c, logK = psk_constel(K), int(round(np.log2(K)))
        k = np.array([int(np.flatnonzero(c == v)[0]) for v in beta.reshape(L, M)[rows, idx]])
        gray = k ^ (k >> 1)
        bits = np.concatenate([bits, ((gray[:, None] >> np.arange(logK)[::-1]) & 1).astype(bool)], axis=1)
    return bits.ravel()


def awgn_channel(x, awgn_var, rand_seed):
    """sparc_sim.py:179-204: variance awgn_var, split over the two dimensions of a complex input."""
    rng = np.random.RandomState(rand_seed)
    if np.iscomplexobj(x):
        return x + np.sqrt(awgn_var / 2) * (rng.randn(x.size) + 1j * rng.randn(x.size))
    return x + np.sqrt(awgn_var) * rng.randn(x.size)


# ------------------------------------------------------------------------ encoder
def normalise_code_params(cp):
    """Defaults of check_code_params (sparc.py:92-102) without its type asserts."""
    cp = dict(cp)
    for k in ("complex", "modulated", "power_allocated", "spatially_coupled"):
        cp.setdefault(k, False)
    if cp["modulated"] and not cp["complex"]:
        assert cp["K"] == 2, "Real-modulated SPARCs requires K=2"
    return cp


def modulation(cp):
    return cp["K"] if cp.get("modulated") else 1


def sparc_encode(code_params, awgn_var, rand_seed):
    """sparc.py:17-53.  Returns bits_in, beta0, x, design (instead of Ab, Az), n."""
    cp = normalise_code_params(code_params)
    L, M, K = cp["L"], cp["M"], modulation(cp)
    bits_in = rnd_bin_arr(int(round(L * np.log2(K * M))), rand_seed)
    beta0 = bin_arr_2_msg_vector(bits_in, M, K)
    W = create_base_matrix(**dict(cp, awgn_var=awgn_var))
    n, _ = code_dims(cp, W)
    design = Design(W, L, M, n, rand_seed, cp["complex"])
    return bits_in, beta0, design.Ab(beta0), design, n


# ------------------------------------------------------------------ AMP decoder
def sparc_amp(y, code_params, decode_params, awgn_var, design, beta0=None, trace=None):
    """sparc.py:883-999 (K=1, real K=2).  `trace`, if a list, receives one dict per
    iteration with the quantities the parity tests compare (tau, phi, psi, s, beta)."""
    cp = normalise_code_params(code_params)
    L, M, n, K = cp["L"], cp["M"], design.n, modulation(cp)
    W = create_base_matrix(**dict(cp, awgn_var=awgn_var))
    t_max = decode_params["t_max"]
    rtol = decode_params.get("rtol", 1e-6)
    method = decode_params.get("phi_est_method", 1)
    if beta0 is None:
        beta0 = np.zeros(L * M)
    beta = np.zeros(L * M) if K in (1, 2) else np.zeros(L * M, dtype=complex)
    z = y
    atol = 2 * np.finfo(float).resolution
    if W.ndim == 0:
        gamma = W
        nmse = np.ones(t_max)
    else:
        if W.ndim == 2:
            Lr = W.shape[0]
            Mr = n // Lr
        Lc = W.shape[-1]
        Mc = L * M // Lc
        gamma = np.dot(W, np.ones(Lc)) / Lc
        nmse = np.ones((t_max, Lc))
    for t in range(t_max - 1):
        if t > 0:
            psi_prev, phi_prev = np.copy(psi), np.copy(phi)
            gamma = W * psi if W.ndim == 0 else np.dot(W, psi) / Lc
            b = gamma / phi_prev
            z = y - design.Ab(beta) + (b * z if W.ndim != 2 else b.repeat(Mr) * z)
        if method == 1:
            phi = awgn_var + gamma
        else:
            phi = (np.abs(z) ** 2).mean() if W.ndim != 2 else (np.abs(z) ** 2).reshape(Lr, -1).mean(axis=1)
        if W.ndim == 0:
            tau = (L * phi / n) / W
            tau_use, phi_use = tau, phi
        elif W.ndim == 1:
            tau = (L * phi / n) / W
            tau_use, phi_use = tau.repeat(Mc), phi
        else:
            tau = (L / Mr) / np.dot(W.T, 1 / phi)
            tau_use, phi_use = tau.repeat(Mc), phi.repeat(Mr)
        s = beta + tau_use * design.Az(z / phi_use)
        beta = mmse_estimator(s, tau_use, M, K)
        if W.ndim == 0:
            psi = 1 - (np.abs(beta) ** 2).sum() / L
            nmse[t + 1] = (np.abs(beta - beta0) ** 2).sum() / L
        else:
            psi = 1 - (np.abs(beta) ** 2).reshape(Lc, -1).sum(axis=1) / (L / Lc)
            nmse[t + 1] = (np.abs(beta - beta0) ** 2).reshape(Lc, -1).sum(axis=1) / (L / Lc)
        if trace is not None:
            trace.append(dict(tau=np.copy(tau), phi=np.copy(phi), psi=np.copy(psi), s=s.copy(), beta=beta.copy(), z=np.copy(z)))
        if t > 0 and np.allclose(psi, psi_prev, rtol, atol=atol):
            nmse[t:] = nmse[t]
            break
    t_final = t + 1
    return map_estimator(s, M, K), t_final, nmse, psi


def sparc_decode(y, code_params, decode_params, awgn_var, design, beta0=None, trace=None):
    """sparc.py:55-74."""
    dp = dict(decode_params)
    dp.setdefault("rtol", 1e-6)
    dp.setdefault("phi_est_method", 1)
    beta, t_final, nmse, psi = sparc_amp(y, code_params, dp, awgn_var, design, beta0, trace)
    expect_err = psi.mean() >= 0.001
    K = code_params["K"] if code_params.get("modulated") else 1
    return msg_vector_2_bin_arr(beta, code_params["M"], K), beta, t_final, nmse, expect_err

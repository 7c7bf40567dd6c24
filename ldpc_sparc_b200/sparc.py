"""Drop-in for the reference's ``sparc_public/sparc.py`` (real, unmodulated SPARCs).

Same function names, argument meaning, dict semantics (``check_code_params``
normalises ``code_params`` IN PLACE, ``sparc_encode`` adds ``n`` / ``R_actual``),
assertion-style validation and return tuples as the reference (sparc.py:17-74),
with the design-matrix operators and the AMP decoder running on the GPU through
libldpcsparc_b200.so:

* ``sparc_encode`` returns ``Ab`` / ``Az`` callables that still work as
  ``Ab(x)`` / ``Az(y)`` on host arrays and carry the device-resident design, so
  ``sparc_decode`` (which receives them from the encoder, sparc_sim.py:21-24)
  runs fully on the GPU;
* ``sparc_decode`` with ``y`` of shape ``[n]`` behaves like the reference; with
  ``[F, n]`` (numpy or CUDA tensor) it decodes F frames that share one design and
  returns batched outputs (``beta`` is then the ``[F, L]`` array of decided
  section indices instead of F dense one-hot vectors).

Every SPARC type of the reference is built: real designs (sub-sampled DCT), unmodulated
or 2-PSK modulated, and complex designs (``complex=True``: sub-sampled DFT, sparc.py:593-646)
with K-PSK modulated sections for any power-of-two K (estimators :402-512).  In batched
outputs a decided section is coded as the symbol ``location + M * k``, k the index into
``psk_constel(K)`` (K = 2: k = 1 is the value -1).
"""
import ctypes as ct
from copy import copy

import numpy as np

from . import _lib

DEFAULT_PRECISION = "fp32"


def is_power_of_2(x):
    return (x > 0) and ((x & (x - 1)) == 0)


# ------------------------------------------------------------ parameter checks
_FLAGS = ('complex', 'modulated', 'power_allocated', 'spatially_coupled')
_FLOAT = (float, np.float64)


def _need(code_params, keys, kept):
    """Copies the required keys into `kept`; same exception text as the reference."""
    if any(k not in code_params for k in keys):
        raise Exception('Need code parameters {}.'.format(list(keys)))
    for k in keys:
        kept[k] = copy(code_params[k])


def check_code_params(code_params):
    """Validates ``code_params`` and normalises it IN PLACE: flags default to False, only
    the keys the selected SPARC type needs survive (semantics of reference sparc.py:77-149:
    exact Python types, AssertionError / Exception on violation, clear() + update())."""
    kept = {}
    for flag in _FLAGS:
        value = code_params.setdefault(flag, False)
        assert type(value) == bool, "'{}' must be boolean".format(flag)
        kept[flag] = copy(value)

    _need(code_params, ('P', 'R', 'L', 'M'), kept)
    P, R, L, M = (code_params[k] for k in ('P', 'R', 'L', 'M'))
    assert type(P) in _FLOAT and P > 0
    assert type(R) in _FLOAT and R > 0
    assert type(L) == int and L > 0
    assert type(M) == int and M > 0 and is_power_of_2(M)

    if kept['modulated']:                       # K-PSK modulation of the non-zero values
        _need(code_params, ('K',), kept)
        K = code_params['K']
        assert type(K) == int and K > 1 and is_power_of_2(K)
        assert kept['complex'] or K == 2, 'Real-modulated SPARCs requires K=2'
    if kept['power_allocated']:                 # iterative power allocation over B blocks
        _need(code_params, ('B', 'R_PA_ratio'), kept)
        B, ratio = code_params['B'], code_params['R_PA_ratio']
        assert type(B) == int and B > 1
        assert L % B == 0, 'B must divide L'
        assert type(ratio) in _FLOAT
        assert ratio >= 0
    if kept['spatially_coupled']:               # (omega, Lambda) band base matrix
        _need(code_params, ('omega', 'Lambda'), kept)
        omega, Lambda = code_params['omega'], code_params['Lambda']
        assert type(omega) == int and omega > 1
        assert type(Lambda) == int and Lambda >= (2 * omega - 1)
        assert L % Lambda == 0, 'Lambda must divide L'
        if kept['power_allocated']:
            assert L % (Lambda * B) == 0, 'Lambda*B must divide L'

    code_params.clear()
    code_params.update(kept)


def check_decode_params(decode_params):
    """``t_max`` is required; ``rtol`` (1e-6) and ``phi_est_method`` (1) get their defaults
    in place (reference sparc.py:151-170)."""
    if 't_max' not in decode_params:
        raise Exception('Need decode parameters {}.'.format(['t_max']))
    decode_params.setdefault('rtol', 1e-6)
    decode_params.setdefault('phi_est_method', 1)
    t_max, rtol = decode_params['t_max'], decode_params['rtol']
    assert type(t_max) == int and t_max > 1
    assert type(rtol) == float and 0 < rtol < 1
    assert decode_params['phi_est_method'] in (1, 2)


def _require_real_unmodulated(code_params):
    """Kept for callers of earlier versions: every SPARC type is built now; a real design still
    needs K = 2 when modulated (sparc.py:117), and the kernels hold at most 64 constellation points."""
    if code_params.get('modulated'):
        assert code_params.get('complex') or code_params.get('K') == 2, 'Real-modulated SPARCs requires K=2'
        if code_params['K'] > 64:
            raise NotImplementedError("K-PSK with K > 64 is not built")


def _modulation(code_params):
    return code_params['K'] if code_params.get('modulated') else 1


# ------------------------------------------------------------ PSK helpers (host)
def bin2gray(num):
    """Binary -> Gray code (reference sparc.py:206-211)."""
    return num ^ (num >> 1)


def gray2bin(num):
    """Gray -> binary code (reference sparc.py:214-223)."""
    for shift in (1, 2, 4, 8, 16, 32):      # prefix XOR by doubling (ints and integer arrays)
        num = num ^ (num >> shift)
    return num


def psk_constel(K):
    """K-PSK constellation symbols (reference sparc.py:225-239)."""
    assert type(K) == int and K > 1 and is_power_of_2(K)
    if K == 2:
        return np.array([1, -1])
    if K == 4:
        return np.array([1 + 0j, 0 + 1j, -1 + 0j, 0 - 1j])
    theta = 2 * np.pi * np.arange(K) / K
    return np.cos(theta) + 1J * np.sin(theta)


def psk_mod(bin_arr, K):
    """Gray-coded K-PSK modulation of ``L*log2(K)`` bits into L symbols (reference sparc.py:241-269)."""
    assert type(K) == int and K > 1 and is_power_of_2(K)
    assert bin_arr.dtype == 'bool'
    c = psk_constel(K)
    logK = int(round(np.log2(K)))
    assert bin_arr.size % logK == 0
    words = bin_arr.reshape(-1, logK).astype(np.int64).dot(1 << np.arange(logK)[::-1])
    symbols = c[[gray2bin(int(w)) for w in words]]
    return symbols[0] if words.size == 1 else symbols


def psk_demod(symbols, K):
    """Inverse of psk_mod (reference sparc.py:271-299)."""
    assert type(K) == int and K > 1 and is_power_of_2(K)
    c = psk_constel(K)
    logK = int(round(np.log2(K)))
    out = []
    for sym in np.atleast_1d(symbols):
        hit = np.flatnonzero(c == sym)
        assert hit.size == 1, 'not a constellation symbol'
        out.append(int_2_bin_arr(bin2gray(int(hit[0])), logK))
    return np.concatenate(out)


# ---------------------------------------------------------------- bit helpers
def rnd_bin_arr(k, rand_seed):
    """Reference sparc.py:174-180."""
    assert type(k) == int
    rng = np.random.RandomState(rand_seed)
    return rng.randint(2, size=k, dtype='bool')


def bin_arr_2_int(bin_array):
    """Reference sparc.py:182-189."""
    assert bin_array.dtype == 'bool'
    k = bin_array.size
    assert 0 < k < 64
    return bin_array.dot(1 << np.arange(k)[::-1])


def int_2_bin_arr(integer, arr_length):
    """Reference sparc.py:191-197 with its numpy-1.x meaning ('0' -> False)."""
    assert integer >= 0
    return np.array([c == '1' for c in np.binary_repr(integer, arr_length)], dtype=bool)


def section_indices(bin_arr, M, K=1):
    """Section symbol ``location + M * k`` of every bit group: logM location bits (MSB first)
    then logK value bits, the Gray code of the constellation index k (psk_mod, sparc.py:241-269).
    ``[..., L*(logM+logK)] -> [..., L]``."""
    logM, logK = int(round(np.log2(M))), int(round(np.log2(K)))
    sec = logM + logK
    b = np.asarray(bin_arr).astype(np.int64)
    assert b.shape[-1] % sec == 0
    b = b.reshape(b.shape[:-1] + (b.shape[-1] // sec, sec))
    idx = b[..., :logM].dot(1 << np.arange(logM)[::-1])
    if K == 1:
        return idx
    return idx + M * gray2bin(b[..., logM:].dot(1 << np.arange(logK)[::-1]))


def indices_2_bin_arr(idx, M, K=1):
    """Inverse of section_indices; ``[..., L] -> [..., L*(logM+logK)]`` bool."""
    logM, logK = int(round(np.log2(M))), int(round(np.log2(K)))
    idx = np.asarray(idx)
    bits = ((idx % M)[..., None] >> np.arange(logM)[::-1]) & 1
    if K > 1:
        bits = np.concatenate([bits, (bin2gray(idx // M)[..., None] >> np.arange(logK)[::-1]) & 1], axis=-1)
    return bits.reshape(idx.shape[:-1] + (idx.shape[-1] * (logM + logK),)).astype(bool)


def symbols_2_msg_vector(sym, M, K=1):
    """Dense message vector(s) of section symbols ``location + M * k``; ``[..., L] -> [..., L*M]``
    (real for K <= 2, complex otherwise, as the reference's vectors)."""
    sym = np.asarray(sym)
    vals = np.ones(sym.shape) if K == 1 else psk_constel(K)[sym // M]
    out = np.zeros(sym.shape + (M,), dtype=vals.dtype if K > 2 else float)
    np.put_along_axis(out, (sym % M)[..., None], vals[..., None], axis=-1)
    return out.reshape(sym.shape[:-1] + (-1,))


def msg_vector_2_symbols(msg_vector, M, K=1):
    """Inverse of symbols_2_msg_vector for vectors with one constellation point per section."""
    rows = np.asarray(msg_vector).reshape(msg_vector.shape[:-1] + (-1, M))
    loc = np.abs(rows).argmax(axis=-1)
    if K == 1:
        return loc
    vals = np.take_along_axis(rows, loc[..., None], axis=-1)[..., 0]
    k = (vals[..., None] * psk_constel(K).conj()).real.argmax(axis=-1)
    return loc + M * k


def rnd_msg_vector(L, M, rand_seed, K=1):
    """Reference sparc.py:303-328: locations first, then values, from one stream."""
    rng = np.random.RandomState(rand_seed)
    assert type(M) == int and M > 0 and is_power_of_2(M)
    msg_vector = np.zeros((L, M)) if K in (1, 2) else np.zeros((L, M), dtype=complex)
    idxs = rng.randint(0, M, L)
    vals = np.ones(L) if K == 1 else psk_constel(K)[rng.randint(0, K, L)]
    msg_vector[(np.arange(L), idxs)] = vals
    return msg_vector.ravel()


def bin_arr_2_msg_vector(bin_arr, M, K=1):
    """Reference sparc.py:330-364: per section logM location bits, then logK Gray-coded value bits."""
    assert type(M) == int and M > 0 and is_power_of_2(M)
    return symbols_2_msg_vector(section_indices(bin_arr, M, K), M, K)


def msg_vector_2_bin_arr(msg_vector, M, K=1):
    """Reference sparc.py:366-400."""
    assert type(msg_vector) == np.ndarray
    assert type(M) == int and M > 0 and is_power_of_2(M)
    assert msg_vector.size % M == 0
    L = msg_vector.size // M
    rows = msg_vector.reshape(L, M)
    idxs1, idxs2 = np.nonzero(rows)
    assert np.array_equal(idxs1, np.arange(L))   # exactly 1 nonzero in each row
    if K > 1:
        hit = rows[idxs1, idxs2][:, None] == psk_constel(K)[None, :]     # exact match, as psk_demod (:271-299)
        assert np.all(hit.sum(axis=1) == 1), 'not a constellation symbol'
        idxs2 = idxs2 + M * hit.argmax(axis=1)
    return indices_2_bin_arr(idxs2, M, K)


def msg_vector_map_estimator(s, M, K=1):
    """Reference sparc.py:467-512 (K = 1: one-hot of the per-section argmax; K = 2: the
    entry of largest magnitude, set to its sign; K > 2: the (location, symbol) pair with the
    largest Re(s conj(c_k)), which is what the reference's K = 4 angle rounding and K > 4
    search both select).  (Output formatting on the host; inside ``sparc_amp`` the decision
    is taken on the GPU.)"""
    if K > 2:
        sm = np.asarray(s).reshape(-1, M)
        corr = (sm[:, :, None] * psk_constel(K).conj()[None, None, :]).real.reshape(sm.shape[0], -1)
        best = corr.argmax(axis=1)
        return symbols_2_msg_vector(best // K + M * (best % K), M, K)
    if K == 2:
        s2 = np.asarray(s).real.reshape(-1, M)
        rows = np.arange(s2.shape[0])
        idx = np.abs(s2).argmax(axis=1)
        beta = np.zeros_like(s2, dtype=float)
        beta[rows, idx] = np.sign(s2[rows, idx])
        return beta.ravel()
    assert type(s) == np.ndarray
    assert s.size % M == 0
    L = s.size // M
    beta = np.zeros((L, M))
    beta[np.arange(L), s.real.reshape(L, -1).argmax(axis=1)] = 1
    return beta.ravel()


# ------------------------------------------------------------- base matrix W
def pa_iterative(P, sigmaSqr, B, R_PA):
    """Iterative power allocation over B blocks (reference sparc.py:516-533): block b gets
    2 ln2 R_PA (sigma^2 + P - mean(Q)) while that exceeds an even split of the remaining
    power, after which the remainder is spread evenly; the result is renormalised to mean P."""
    Q = np.zeros(B)
    b = 0
    while b < B:
        greedy = 2 * np.log(2) * R_PA * (sigmaSqr + P - Q.mean())
        even = (B * P - Q.sum()) / (B - b)
        if greedy <= even:
            Q[b:] = even
            break
        Q[b] = greedy
        b += 1
    return Q / (Q.mean() / P)


def sc_basic(Q, omega, Lambda):
    """(omega, Lambda) spatially coupled base matrix with uncoupled entry / vector Q
    (reference sparc.py:535-568): column block c carries Q * Lr / omega in rows c..c+omega-1."""
    assert type(Q) == np.ndarray
    if Q.ndim > 1:
        raise Exception('Something wrong with Q')
    Lr = Lambda + omega - 1
    r, c = np.arange(Lr)[:, None], np.arange(Lambda)[None, :]
    band = ((r >= c) & (r < c + omega)).astype(float)              # [Lr, Lambda]
    W = band * (Q * Lr / omega) if Q.ndim == 0 else np.kron(band, (Q * Lr / omega)[None, :])
    assert np.isclose(W.mean(), np.mean(Q)), "Average base matrix values must equal P"
    return W


def create_base_matrix(P, power_allocated=False, spatially_coupled=False, **kwargs):
    """Base entry / vector / matrix W of the design (reference sparc.py:570-589)."""
    Q = np.array(P)
    if power_allocated:
        Q = pa_iterative(P, kwargs['awgn_var'], kwargs['B'], kwargs['R'] * kwargs['R_PA_ratio'])
    return sc_basic(Q, kwargs['omega'], kwargs['Lambda']) if spatially_coupled else Q


# ----------------------------------------------------------- design + operators
class Design:
    """Sub-sampled DCT (real) or DFT (``csparc``) design matrix: host description + device handle.

    Orderings are drawn exactly as the reference's generate_ordering
    (sparc.py:735-775): one ``RandomState(rand_seed)``, two index arrays (``arange(1, w)``;
    for the DFT design ``arange(w)`` without 0 and w/2) shuffled in place, cumulatively,
    once per non-zero block."""

    def __init__(self, W, L, M, n, rand_seed, csparc=False):
        self.csparc = bool(csparc)
        assert type(W) == np.ndarray
        assert type(L) == int and type(M) == int and type(n) == int
        assert L > 0 and M > 0 and n > 0
        self.W, self.L, self.M, self.n = W, L, M, n
        if W.ndim == 0:
            Wm = W.reshape(1, 1).astype(float)
        elif W.ndim == 1:
            Wm = W.reshape(1, -1).astype(float)
            assert L % W.size == 0
        elif W.ndim == 2:
            Wm = W.astype(float)
            assert L % W.shape[1] == 0
            assert n % W.shape[0] == 0
        else:
            raise Exception('Something wrong with base matrix input W')
        self.Wm = np.ascontiguousarray(Wm)
        self.Lr, self.Lc = Wm.shape
        self.Mr, self.Mc = n // self.Lr, L * M // self.Lc
        if not self.csparc:
            self.w = 2 ** int(np.ceil(np.log2(max(self.Mr + 1, self.Mc + 1))))
            idxs0 = np.arange(1, self.w, dtype=np.uint32)
            idxs1 = np.arange(1, self.w, dtype=np.uint32)
        else:
            self.w = 2 ** int(np.ceil(np.log2(max(self.Mr + 2, self.Mc + 2))))
            idxs0 = np.delete(np.arange(self.w, dtype=np.uint32), [0, self.w // 2])
            idxs1 = np.delete(np.arange(self.w, dtype=np.uint32), [0, self.w // 2])
        rng = np.random.RandomState(rand_seed)
        self.order0 = np.zeros((self.Lr, self.Lc, self.Mr), dtype=np.uint32)
        self.order1 = np.zeros((self.Lr, self.Lc, self.Mc), dtype=np.uint32)
        for r in range(self.Lr):
            for c in range(self.Lc):
                if Wm[r, c] != 0:
                    rng.shuffle(idxs0)
                    rng.shuffle(idxs1)
                    self.order0[r, c] = idxs0[:self.Mr]
                    self.order1[r, c] = idxs1[:self.Mc]
        self._handle = None

    def handle(self):
        if self._handle is None:
            _lib.require_device()
            h = ct.c_void_p()
            create = _lib.lib().lsb_csparc_design_create if self.csparc else _lib.lib().lsb_sparc_design_create
            rc = create(
                self.Wm.ctypes.data_as(_lib.c_f64p), self.Lr, self.Lc, self.L, self.M, self.n,
                self.order0.ctypes.data_as(_lib.c_u32p), self.order1.ctypes.data_as(_lib.c_u32p), ct.byref(h))
            if rc == -1 and "transform size" in _lib.last_error():
                raise NotImplementedError(_lib.last_error())
            _lib.check(rc, "lsb_sparc_design_create")
            self._handle = h
        return self._handle

    def band_handle(self):
        """Explicit (block-banded, fp16 hi / lo) copy of the design for the tensor-core path
        (``precision='tc'``, csrc/sparc_band_tc.cu); built on first use."""
        if self.csparc:
            raise NotImplementedError("precision='tc' handles the real (DCT) design")
        if getattr(self, "_band", None) is None:
            _lib.require_device()
            h = ct.c_void_p()
            _lib.check(_lib.lib().lsb_band_design_create(
                self.Wm.ctypes.data_as(_lib.c_f64p), self.Lr, self.Lc, self.L, self.M, self.n,
                self.order0.ctypes.data_as(_lib.c_u32p), self.order1.ctypes.data_as(_lib.c_u32p), ct.byref(h)),
                "lsb_band_design_create")
            self._band = h
        return self._band

    def __del__(self):
        try:
            if getattr(self, "_band", None) is not None:
                _lib.lib().lsb_band_design_destroy(self._band)
                self._band = None
        except Exception:
            pass
        try:
            if getattr(self, "_handle", None) is not None:
                destroy = _lib.lib().lsb_csparc_design_destroy if self.csparc else _lib.lib().lsb_sparc_design_destroy
                destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def apply(self, vec, transpose, precision="fp64"):
        """A @ vec (transpose=False, length L*M -> n) or A.T @ vec, on the GPU.
        ``vec`` is ``[len]`` or ``[F, len]``; numpy in -> float64 numpy out, CUDA
        tensor in -> tensor out."""
        import torch
        if self.csparc:
            return self._apply_complex(vec, transpose, precision)
        n_in = self.n if transpose else self.L * self.M
        n_out = self.L * self.M if transpose else self.n
        is_tensor = isinstance(vec, torch.Tensor)
        dt = torch.float64 if precision == "fp64" else torch.float32
        if is_tensor:
            x = vec.to(dt)
        else:
            _lib.require_device()
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(vec, dtype=np.float64))).to("cuda").to(dt)
        single = x.dim() == 1
        assert x.shape[-1] == n_in, "x must be %d long" % n_in
        x = x.reshape(-1, n_in).contiguous()
        out = torch.empty((x.shape[0], n_out), dtype=dt, device=x.device)
        L_ = _lib.lib()
        fn = {(False, "fp32"): L_.lsb_sparc_ab_dev_f32, (False, "fp64"): L_.lsb_sparc_ab_dev_f64,
              (True, "fp32"): L_.lsb_sparc_az_dev_f32, (True, "fp64"): L_.lsb_sparc_az_dev_f64,
              (False, "tc"): L_.lsb_band_ab_dev_f32, (True, "tc"): L_.lsb_band_az_dev_f32}[(transpose, precision)]
        with torch.cuda.device(x.device):
            _lib.check(fn(self.band_handle() if precision == "tc" else self.handle(), x.data_ptr(), out.data_ptr(), x.shape[0],
                          torch.cuda.current_stream().cuda_stream), "lsb_sparc_ab/az")
        if single:
            out = out[0]
        return out if is_tensor else out.cpu().numpy().astype(np.float64)


    def _apply_complex(self, vec, transpose, precision):
        """DFT design: A @ vec or A^H @ vec on complex vectors (real input is promoted)."""
        import torch
        n_in = self.n if transpose else self.L * self.M
        n_out = self.L * self.M if transpose else self.n
        is_tensor = isinstance(vec, torch.Tensor)
        dt = torch.complex128 if precision == "fp64" else torch.complex64
        if is_tensor:
            x = vec.to(dt)
        else:
            _lib.require_device()
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(vec, dtype=np.complex128))).to("cuda").to(dt)
        single = x.dim() == 1
        assert x.shape[-1] == n_in, "x must be %d long" % n_in
        x = x.reshape(-1, n_in).contiguous()
        out = torch.empty((x.shape[0], n_out), dtype=dt, device=x.device)
        L_ = _lib.lib()
        fn = {(False, "fp32"): L_.lsb_csparc_ab_dev_f32, (False, "fp64"): L_.lsb_csparc_ab_dev_f64,
              (True, "fp32"): L_.lsb_csparc_az_dev_f32, (True, "fp64"): L_.lsb_csparc_az_dev_f64}[(transpose, precision)]
        with torch.cuda.device(x.device):
            _lib.check(fn(self.handle(), x.data_ptr(), out.data_ptr(), x.shape[0],
                          torch.cuda.current_stream().cuda_stream), "lsb_csparc_ab/az")
        if single:
            out = out[0]
        return out if is_tensor else out.cpu().numpy().astype(np.complex128)


class _Operator:
    """Callable returned by sparc_transforms: ``Ab(x)`` / ``Az(y)`` (sparc.py:786-875)."""

    def __init__(self, design, transpose):
        self.design, self.transpose = design, transpose

    def __call__(self, v, precision="fp64"):
        return self.design.apply(v, self.transpose, precision)


def sparc_transforms(W, L, M, n, rand_seed, csparc=False):
    """Returns ``Ab(x)`` (A times x, x of length L*M) and ``Az(y)`` (A transpose
    times y -- conjugate transpose for ``csparc`` --, y of length n) for the sub-sampled DCT
    design or, with ``csparc``, the sub-sampled DFT design (reference sparc.py:703-880)."""
    design = Design(W, L, M, n, rand_seed, csparc)
    return _Operator(design, False), _Operator(design, True)


# -------------------------------------------------------------------- encoder
def sparc_encode(code_params, awgn_var, rand_seed):
    """Encode a random message to a SPARC codeword (reference sparc.py:17-53).
    Returns ``bits_in, beta0, x, Ab, Az``."""
    check_code_params(code_params)
    _require_real_unmodulated(code_params)
    R, L, M = map(code_params.get, ['R', 'L', 'M'])
    K = _modulation(code_params)
    bit_len = int(round(L * np.log2(K * M)))
    bits_in = rnd_bin_arr(bit_len, rand_seed)
    beta0 = bin_arr_2_msg_vector(bits_in, M, K)

    tmp = code_params.copy()
    tmp.update({'awgn_var': awgn_var})
    W = create_base_matrix(**tmp)

    n = int(round(bit_len / R))
    if W.ndim == 2:
        Lr, _ = W.shape
        Mr = int(round(n / Lr))
        n = Mr * Lr
    R_actual = bit_len / n
    code_params.update({'n': n, 'R_actual': R_actual})

    Ab, Az = sparc_transforms(W, L, M, n, rand_seed, code_params['complex'])
    x = Ab(beta0)
    return bits_in, beta0, x, Ab, Az


# -------------------------------------------------------------------- decoder
def sparc_amp(y, code_params, decode_params, awgn_var, rand_seed, beta0, Ab=None, Az=None,
              precision=None, return_soft=False, return_trace=False):
    """AMP decoder (reference sparc.py:883-999).  Returns ``beta, t_final, nmse, psi``
    (single frame: dense one-hot ``beta``; batch: ``[F, L]`` section indices).
    With ``return_soft`` / ``return_trace`` a dict of extra arrays (soft beta, s,
    per-iteration tau and psi) is appended -- used by the parity tests."""
    import torch
    _require_real_unmodulated(code_params)
    P, R, L, M, n = map(code_params.get, ['P', 'R', 'L', 'M', 'n'])
    K = _modulation(code_params)
    tmp = code_params.copy()
    tmp.update({'awgn_var': awgn_var})
    W = create_base_matrix(**tmp)
    assert 0 <= W.ndim <= 2
    t_max, rtol, phi_est_method = map(decode_params.get, ['t_max', 'rtol', 'phi_est_method'])
    assert phi_est_method == 1 or phi_est_method == 2

    if (Ab is None) or (Az is None):
        Ab, Az = sparc_transforms(W, L, M, n, rand_seed, code_params['complex'])
    if not isinstance(Ab, _Operator):
        raise TypeError("Ab/Az must come from this module's sparc_transforms / sparc_encode "
                        "(arbitrary Python closures cannot run on the GPU)")
    design = Ab.design
    cplx = design.csparc
    precision = precision or DEFAULT_PRECISION
    if cplx:
        dt = torch.complex128 if precision == "fp64" else torch.complex64
    else:
        dt = torch.float64 if precision == "fp64" else torch.float32

    is_tensor = isinstance(y, torch.Tensor)
    if is_tensor:
        yt = y.to(dt)
    else:
        _lib.require_device()
        host = np.asarray(y, dtype=np.complex128 if cplx else np.float64)
        yt = torch.from_numpy(np.ascontiguousarray(host)).to("cuda").to(dt)
    single = yt.dim() == 1
    assert yt.shape[-1] == n
    yt = yt.reshape(-1, n).contiguous()
    F, dev = yt.shape[0], yt.device
    Lc = design.Lc

    idx0 = None
    if isinstance(beta0, torch.Tensor) and beta0.is_cuda and tuple(beta0.shape) == (F, L) and M != 1:
        idx0 = beta0.to(dev, torch.int32).contiguous()        # section indices already on the device
    elif beta0 is not None:
        b0 = beta0.cpu().numpy() if isinstance(beta0, torch.Tensor) else np.asarray(beta0)
        if b0.shape[-1] == L * M:                       # dense message vector(s), as the reference passes
            b0 = msg_vector_2_symbols(b0.reshape(-1, L * M), M, K)
        b0 = np.broadcast_to(b0.reshape(-1, L), (F, L)) if b0.reshape(-1, L).shape[0] in (1, F) else None
        if b0 is not None:
            idx0 = torch.from_numpy(np.array(b0, dtype=np.int32)).to(dev)

    idx = torch.empty((F, L), dtype=torch.int32, device=dev)
    t_final = torch.empty(F, dtype=torch.int32, device=dev)
    psi = torch.empty((F, Lc), dtype=torch.float64, device=dev)
    nmse = torch.ones((F, t_max, Lc), dtype=torch.float32, device=dev) if idx0 is not None else None
    beta_soft = torch.empty((F, L * M), dtype=dt, device=dev) if return_soft else None
    s_last = torch.empty((F, L * M), dtype=dt, device=dev) if return_soft else None
    tau_tr = torch.zeros((F, t_max, Lc), dtype=torch.float64, device=dev) if return_trace else None
    psi_tr = torch.zeros((F, t_max, Lc), dtype=torch.float64, device=dev) if return_trace else None

    # atol of the early-stop test: the reference hard-codes 2*finfo(float).resolution = 2e-15
    # (sparc.py:916).  decode_params['atol'] (an extension, not validated by
    # check_decode_params) lets an fp32 run stop at its own noise floor, e.g. 1e-7.
    atol = float(decode_params.get('atol', 2 * np.finfo(float).resolution))
    params = _lib.AmpParams(float(awgn_var), float(rtol), atol, int(t_max), int(phi_est_method), int(K), 0)
    ptr = lambda t: t.data_ptr() if t is not None else None
    out = _lib.AmpOut(ptr(idx), ptr(t_final), ptr(psi), ptr(nmse), ptr(beta_soft), ptr(s_last), ptr(tau_tr), ptr(psi_tr))
    if precision == "tc":
        # explicit block-banded design, A beta / A^T z as tcgen05 GEMMs over the frame batch (csrc/sparc_band_tc.cu)
        if cplx or K != 1:
            raise NotImplementedError("precision='tc' handles real, unmodulated SPARCs")
        fn, handle = _lib.lib().lsb_band_amp_dev_f32, design.band_handle()
    else:
        fn = getattr(_lib.lib(), "lsb_%s_amp_dev_%s" % ("csparc" if cplx else "sparc", "f64" if precision == "fp64" else "f32"))
        handle = design.handle()
    with torch.cuda.device(dev):
        _lib.check(fn(handle, yt.data_ptr(), F, ct.byref(params), ptr(idx0), ct.byref(out),
                      torch.cuda.current_stream().cuda_stream), "lsb_sparc_amp_dev")

    extra = {}
    if return_soft:
        extra.update(beta_soft=beta_soft, s=s_last)
    if return_trace:
        extra.update(tau=tau_tr, psi=psi_tr)

    if is_tensor and not single:
        res = (idx, t_final, nmse, psi)
        return res + (extra,) if extra else res

    t_np = t_final.cpu().numpy()
    psi_np = psi.cpu().numpy()
    idx_np = idx.cpu().numpy()
    if nmse is not None:
        nm = nmse.cpu().numpy().astype(np.float64)   # the kernel applies nmse[t:] = nmse[t] on early stop (sparc.py:985)
    else:
        nm = np.ones((F, t_max, Lc))
    if W.ndim == 0:
        nm = nm[:, :, 0]
    extra = {k: v.cpu().numpy() for k, v in extra.items()}
    if single:
        beta = symbols_2_msg_vector(idx_np[0], M, K).reshape(L, M)
        psi_out = psi_np[0, 0] if W.ndim == 0 else psi_np[0]
        res = (beta.ravel(), int(t_np[0]), nm[0], psi_out)
        extra = {k: v[0] for k, v in extra.items()}
    else:
        res = (idx_np, t_np, nm, psi_np[:, 0] if W.ndim == 0 else psi_np)
    return res + (extra,) if extra else res


def sparc_decode_host(y, code_params, decode_params, awgn_var, Ab, out=None, precision="fp32"):
    """Batched fp32 AMP decode through the host-pointer C-ABI entry ``lsb_sparc_amp_host_f32``: the call a
    Monte-Carlo loop over the reference's ``sparc_decode`` (sparc.py:55-74, sparc_sim.py:21-24) makes, for
    ``F`` frames at once.  ``y``: float32 ``[F, n]`` in host memory (a numpy array, or a pinned CPU torch
    tensor for asynchronous copies).  Returns ``(idx [F, L] int32, t_final [F] int32, psi [F, Lc] float64)``
    as numpy arrays (or fills the pinned torch tensors passed in ``out``); ``expect_err = psi.mean(1) >= 1e-3``.
    ``precision='tc'``: the explicit-design tensor-core path (``lsb_band_amp_host_f32``)."""
    import torch
    check_decode_params(decode_params)
    _require_real_unmodulated(code_params)
    if not isinstance(Ab, _Operator) or Ab.design.csparc:
        raise TypeError("Ab must come from this module's sparc_transforms / sparc_encode (real design)")
    design = Ab.design
    L, n, Lc = code_params['L'], code_params['n'], design.Lc
    if isinstance(y, torch.Tensor):
        assert y.device.type == "cpu" and y.dtype == torch.float32 and y.is_contiguous()
        F, y_ptr = y.reshape(-1, n).shape[0], y.data_ptr()
    else:
        y = np.ascontiguousarray(y, dtype=np.float32).reshape(-1, n)
        F, y_ptr = y.shape[0], y.ctypes.data
    _lib.require_device()
    if out is None:
        idx, t_final, psi = np.empty((F, L), np.int32), np.empty(F, np.int32), np.empty((F, Lc), np.float64)
        ptrs = (idx.ctypes.data, t_final.ctypes.data, psi.ctypes.data)
    else:
        idx, t_final, psi = out
        ptrs = (idx.data_ptr(), t_final.data_ptr(), psi.data_ptr() if psi is not None else None)
    t_max, rtol, phi_est_method = map(decode_params.get, ['t_max', 'rtol', 'phi_est_method'])
    atol = float(decode_params.get('atol', 2 * np.finfo(float).resolution))
    params = _lib.AmpParams(float(awgn_var), float(rtol), atol, int(t_max), int(phi_est_method), int(_modulation(code_params)), 0)
    if precision == "tc":
        _lib.check(_lib.lib().lsb_band_amp_host_f32(design.band_handle(), y_ptr, F, ct.byref(params), *ptrs), "lsb_band_amp_host_f32")
    else:
        _lib.check(_lib.lib().lsb_sparc_amp_host_f32(design.handle(), y_ptr, F, ct.byref(params), *ptrs), "lsb_sparc_amp_host_f32")
    return idx, t_final, psi


def sparc_decode(y, code_params, decode_params, awgn_var, rand_seed, beta0, Ab=None, Az=None, precision=None):
    """Decode SPARC codeword(s) (reference sparc.py:55-74).  Returns
    ``bits_out, beta, t_final, nmse, expect_err``."""
    import torch
    check_decode_params(decode_params)
    beta, t_final, nmse, psi = sparc_amp(y, code_params, decode_params, awgn_var, rand_seed, beta0, Ab, Az,
                                         precision=precision)
    M, K = code_params['M'], _modulation(code_params)
    if isinstance(beta, torch.Tensor):                  # batched device path
        expect_err = psi.mean(dim=1) >= 0.001
        logM = int(round(np.log2(M)))
        shifts = torch.arange(logM - 1, -1, -1, device=beta.device, dtype=torch.int32)
        bits_out = (((beta % M).unsqueeze(-1) >> shifts) & 1)
        if K > 1:                                       # Gray-coded value bits after the location bits
            k = beta // M
            kshifts = torch.arange(int(round(np.log2(K))) - 1, -1, -1, device=beta.device, dtype=torch.int32)
            bits_out = torch.cat([bits_out, (((k ^ (k >> 1)).unsqueeze(-1) >> kshifts) & 1)], dim=-1)
        return bits_out.reshape(beta.shape[0], -1).bool(), beta, t_final, nmse, expect_err
    if beta.ndim == 1 and beta.size == code_params['L'] * M:      # single frame: dense beta
        expect_err = psi.mean() >= 0.001
        return msg_vector_2_bin_arr(beta, M, K), beta, t_final, nmse, expect_err
    psi2 = np.asarray(psi).reshape(len(beta), max(1, np.asarray(psi).size // max(len(beta), 1)))
    return indices_2_bin_arr(beta, M, K), beta, t_final, nmse, psi2.mean(axis=1) >= 0.001

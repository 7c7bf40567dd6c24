"""Drop-in for the reference's ``sparc_sophie/sparc_new.py``: concatenated LDPC-SPARC
with a dense Gaussian design matrix, AMP, and the three AMP+BP decoders.

Same function names, arity and return values as the reference for the working
decoders (sparc_new.py:15-82, 257-282, 472-502, 709-735, 885-912, 975-990,
1040-1066, 1099-1193, 1260-1388); the experimental ``*_test*`` / ``*_loop`` /
``*_posteriors`` variants (out of scope) are not rebuilt.

All decoding arithmetic runs on the GPU (csrc/sparc_dense.cu + the BP kernels).
``y`` may be ``[n]`` (the reference's call) or ``[F, n]`` (F frames sharing the
design ``A``); ``A`` is a NumPy array -- uploaded once and cached -- exactly what
``sparc_ldpc_encode`` returns.  ``precision`` is 'fp64' (parity mode, default, the
reference's float64), 'fp32' (SIMT fp32) or 'tc' (throughput mode: the two products
A beta and A^T z run on the tensor cores -- tcgen05, split-fp16 operands with fp32
accumulation, about 2e-5 relative -- and everything else in fp32).
"""
import ctypes as ct

import numpy as np

from . import _lib
from .ldpc import code

DEFAULT_PRECISION = "fp64"


# ------------------------------------------------------------------ host helpers
def bin_arr_2_int(bin_array):
    """sparc_new.py:1362-1370."""
    assert bin_array.dtype == 'bool'
    k = bin_array.size
    assert 0 < k < 64
    return bin_array.dot(1 << np.arange(k)[::-1])


def int_2_bin_arr(integer, arr_length):
    """sparc_new.py:1372-1378 with its numpy-1.x meaning ('0' -> False)."""
    assert integer >= 0
    return np.array([ch == '1' for ch in np.binary_repr(integer, arr_length)], dtype=bool)


def bit_err_rate(bits_in, bits_out):
    """sparc_new.py:1380-1388."""
    return np.sum(np.asarray(bits_in) != np.asarray(bits_out)) / len(bits_in)


def create_design_matrix(L, M, n, rand_seed):
    """sparc_new.py:1284-1294: i.i.d. N(0, 1/n) entries from default_rng(rand_seed)."""
    rng = np.random.default_rng(rand_seed)
    return rng.normal(loc=0, scale=1 / np.sqrt(n), size=(n, M * L))


def bin_arr_2_msg_vector(bin_arr, M, n, P_l):
    """sparc_new.py:1298-1318."""
    logM = int(np.log2(M))
    assert bin_arr.size % logM == 0
    L = bin_arr.size // logM
    idx = bin_arr.reshape(L, logM).astype(np.int64).dot(1 << np.arange(logM)[::-1])
    msg_vector = np.zeros(L * M)
    msg_vector[np.arange(L) * M + idx] = np.sqrt(n * P_l)
    return msg_vector


def msg_vector_2_bin_arr(msg_vector, M):
    """sparc_new.py:1320-1341."""
    assert type(msg_vector) == np.ndarray
    assert type(M) == int and M > 0
    assert msg_vector.size % M == 0
    logM = int(round(np.log2(M)))
    L = msg_vector.size // M
    idxs1, idxs2 = np.nonzero(msg_vector.reshape(L, M))
    assert np.array_equal(idxs1, np.arange(L))
    return ((idxs2[:, None] >> np.arange(logM)[::-1]) & 1).astype(bool).ravel()


def msg_vector_map_estimator(s, M, sqrt_nP_l):
    """sparc_new.py:1099-1116 (host-side output formatting)."""
    L = s.size // M
    beta = np.zeros((L, M))
    beta[np.arange(L), s.reshape(L, -1).argmax(axis=1)] = sqrt_nP_l
    return beta.ravel()


def S_k_mapping(M):
    """sparc_new.py:1140-1160: S_k[b] = indices whose bit b (b = 0 is the MSB) is 0."""
    logM = int(np.log2(M))
    S_k = [[] for _ in range(logM)]
    for i in range(logM):
        b = logM - 1 - i
        k = 0
        while k < M:
            for j in range(k, k + pow(2, i)):
                S_k[b].append(j)
            k = k + pow(2, i + 1)
    return S_k


def encode_ldpc(user_bits, ldpc_params, mults, unprotected_bit_len):
    """sparc_new.py:1343-1359."""
    c = code(ldpc_params["standard"], ldpc_params["rate"], ldpc_params["z"])
    ldpc_bits = user_bits[unprotected_bit_len:]
    unprotected_bits = user_bits[:unprotected_bit_len].astype(bool)
    ldpc_bits = np.array_split(ldpc_bits, mults)
    for i in range(mults):
        ldpc_bits[i] = c.encode(ldpc_bits[i])
    ldpc_bits = np.concatenate(ldpc_bits).astype(bool)
    return np.concatenate((unprotected_bits, ldpc_bits))


# ------------------------------------------------------------------ device engine
_design_cache = {}


class _Engine:
    """Device-resident design matrix + typed wrappers of the C-ABI dense ops."""

    def __init__(self, A, precision):
        import torch
        _lib.require_device()
        self.torch = torch
        self.precision = precision
        self.tc = precision == "tc"          # tensor-core products (tcgen05, split fp16), fp32 everywhere else
        self.bp_precision = "fp64" if precision == "fp64" else "fp32"
        self.dt = torch.float64 if precision == "fp64" else torch.float32
        self.suf = "f64" if precision == "fp64" else "f32"
        A = np.ascontiguousarray(A, dtype=np.float64)
        self.n, self.ML = A.shape
        h = ct.c_void_p()
        if self.tc:
            _lib.check(_lib.lib().lsb_tc_design_create(A.ctypes.data_as(_lib.c_f64p), self.n, self.ML, ct.byref(h)),
                       "lsb_tc_design_create")
        else:
            _lib.check(_lib.lib().lsb_dense_design_create(A.ctypes.data_as(_lib.c_f64p), self.n, self.ML,
                                                          int(precision == "fp32"), int(precision == "fp64"),
                                                          ct.byref(h)), "lsb_dense_design_create")
        self.h = h
        self.dev = torch.device("cuda", torch.cuda.current_device())

    def __del__(self):
        try:
            (_lib.lib().lsb_tc_design_destroy if self.tc else _lib.lib().lsb_dense_design_destroy)(self.h)
        except Exception:
            pass

    def fn(self, name):
        return getattr(_lib.lib(), "lsb_%s_dev_%s" % (name, self.suf))

    def stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def to_dev(self, v):
        t = self.torch
        if isinstance(v, t.Tensor):
            return v.to(self.dev, self.dt).reshape(-1, v.shape[-1]).contiguous()
        v = np.asarray(v, dtype=np.float64)
        return t.from_numpy(np.ascontiguousarray(v.reshape(-1, v.shape[-1]))).to(self.dev).to(self.dt)

    def empty(self, *shape, dtype=None):
        return self.torch.empty(shape, dtype=dtype or self.dt, device=self.dev)

    def ab(self, beta):
        out = self.empty(beta.shape[0], self.n)
        fn = _lib.lib().lsb_tc_ab_dev_f32 if self.tc else self.fn("dense_ab")
        _lib.check(fn(self.h, beta.data_ptr(), out.data_ptr(), beta.shape[0], self.stream()), "dense_ab")
        return out

    def az(self, z):
        out = self.empty(z.shape[0], self.ML)
        fn = _lib.lib().lsb_tc_az_dev_f32 if self.tc else self.fn("dense_az")
        _lib.check(fn(self.h, z.data_ptr(), out.data_ptr(), z.shape[0], self.stream()), "dense_az")
        return out

    def residual(self, y, ab, z, beta_sq, onsager, tau, P, mode, first):
        p = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self.fn("amp_residual")(y.data_ptr(), p(ab), z.data_ptr(), p(beta_sq), p(onsager), tau.data_ptr(),
                                           y.shape[0], self.n, float(P), mode, int(first), self.stream()), "amp_residual")

    def denoise(self, beta_in, atz, beta_out, s_out, tau, beta_sq, L, M, sq):
        p = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self.fn("amp_denoise")(p(beta_in), atz.data_ptr(), beta_out.data_ptr(), p(s_out), tau.data_ptr(),
                                          beta_sq.data_ptr(), atz.shape[0], L, M, float(sq), self.stream()), "amp_denoise")

    def section_to_bit(self, beta, L, M, sq, want_probs=True, want_llr=True):
        logM = int(np.log2(M))
        F = beta.shape[0]
        probs = self.empty(F, L * logM) if want_probs else None
        llr = self.empty(F, L * logM) if want_llr else None
        p = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self.fn("section_to_bit")(beta.data_ptr(), p(probs), p(llr), F * L, M, float(sq), self.stream()),
                   "section_to_bit")
        return probs, llr

    def bit_to_section(self, app, L, M, sq, beta_sq=None, want_probs=True):
        F = app.shape[0]
        logM = int(np.log2(M))
        probs = self.empty(F, L * logM) if want_probs else None
        beta = self.empty(F, L * M)
        p = lambda t: t.data_ptr() if t is not None else None
        _lib.check(self.fn("bit_to_section")(app.data_ptr(), p(probs), beta.data_ptr(), p(beta_sq), F, L, M, float(sq),
                                             self.stream()), "bit_to_section")
        return probs, beta

    def bp(self, llr, c, num_its):
        """LLRs [F, blocks*N] -> APP, fresh messages per block (sparc_new.py:1172-1179)."""
        F = llr.shape[0]
        app, _ = c.decode(llr.reshape(-1, c.N), num_its, precision=self.bp_precision)
        return app.reshape(F, -1)

    def hard_bits(self, app, c):
        F = app.shape[0]
        blocks = app.shape[1] // c.N
        bits = self.empty(F, blocks * c.K, dtype=self.torch.uint8)
        _lib.check(self.fn("hard_info_bits")(app.data_ptr(), bits.data_ptr(), F * blocks, c.N, c.K, self.stream()),
                   "hard_info_bits")
        return bits

    def onsager(self, beta, vk, vk0, alpha_w, tau, L, M, sq):
        out = self.empty(beta.shape[0], dtype=self.torch.float64)
        _lib.check(self.fn("integrated_onsager")(beta.data_ptr(), vk.data_ptr(), vk0.data_ptr(), alpha_w.data_ptr(),
                                                 tau.data_ptr(), out.data_ptr(), beta.shape[0], L, M, float(sq),
                                                 self.stream()), "integrated_onsager")
        return out


def _engine(A, precision):
    precision = precision or DEFAULT_PRECISION
    if precision not in ("fp32", "fp64", "tc"):
        raise ValueError("precision must be 'fp32', 'fp64' or 'tc'")
    if isinstance(A, _Engine):
        return A
    key = (id(A), A.shape, precision)
    eng = _design_cache.get(key)
    if eng is None or eng._src() is not A:
        import weakref
        if len(_design_cache) >= 4:
            _design_cache.clear()
        eng = _Engine(A, precision)
        try:
            eng._src = weakref.ref(A)
        except TypeError:
            eng._src = lambda: None
        _design_cache[key] = eng
    return eng


def _out(t, single):
    a = t.cpu().numpy()
    if a.dtype == np.float32:
        a = a.astype(np.float64)
    return a[0] if single else a


def _is_single(y):
    return len(y.shape) == 1


# ------------------------------------------------------------------ two half-batches on two streams
TWO_STREAM_MIN_FRAMES = 2048     # measured on B200 (config 3): +12 % at 2048 frames, no gain with halves of 512


def _drive(eng, yd, make_gen):
    """Runs the decoder generator ``make_gen(y_chunk)`` (one ``yield`` per AMP iteration, result
    returned at the end) over the frames of ``yd``.  With tensor-core products and a large batch
    the frames are cut into two halves that advance alternately on two CUDA streams, so the
    memory-bound kernels of one half (denoiser, residual, glue, BP) run under the GEMMs of the
    other.  Frames are independent: the outputs are those of two separate calls."""
    t = eng.torch
    F = yd.shape[0]
    if not eng.tc or F < TWO_STREAM_MIN_FRAMES:
        gen = make_gen(yd)
        while True:
            try:
                next(gen)
            except StopIteration as e:
                return e.value
    if not hasattr(eng, "_streams"):
        eng._streams = [t.cuda.Stream(device=eng.dev), t.cuda.Stream(device=eng.dev)]
    cur = t.cuda.current_stream()
    cut = (F // 2 + 127) // 128 * 128                      # whole GEMM frame tiles in the first half
    halves = [yd[:cut], yd[cut:]]
    gens, results, live = [], [None, None], [True, True]
    for h, st in zip(halves, eng._streams):
        st.wait_stream(cur)
        gens.append(make_gen(h))
    while any(live):
        for i, st in enumerate(eng._streams):
            if live[i]:
                with t.cuda.stream(st):
                    try:
                        next(gens[i])
                    except StopIteration as e:
                        results[i], live[i] = e.value, False
    for st in eng._streams:
        cur.wait_stream(st)
    for r in results:                                      # the halves' outputs are read on the caller's stream from here on
        _record_stream(t, r, cur)
    return _join(t, results[0], results[1])


def _record_stream(t, x, stream):
    if isinstance(x, tuple):
        for y in x:
            _record_stream(t, y, stream)
    elif isinstance(x, t.Tensor) and x.is_cuda:
        x.record_stream(stream)


def _join(t, a, b):
    if isinstance(a, tuple):
        return tuple(_join(t, x, y) for x, y in zip(a, b))
    if isinstance(a, np.ndarray):
        return np.concatenate([a, b], axis=0)
    return t.cat([a, b], dim=0)


# ------------------------------------------------------------------ AMP
def _amp_state(eng, y, L, M):
    F = y.shape[0]
    t = eng.torch
    return dict(beta=t.zeros((F, L * M), dtype=eng.dt, device=eng.dev), z=eng.empty(F, eng.n),
                tau=t.ones(F, dtype=t.float64, device=eng.dev), beta_sq=t.zeros(F, dtype=t.float64, device=eng.dev))


def _amp_iteration(eng, y, st, P, L, M, sq, first, s_out=None):
    """One AMP iteration (sparc_new.py:975-990); st['beta'] <- eta(s, tau^2)."""
    if first:
        eng.residual(y, None, st["z"], None, None, st["tau"], P, 0, True)
    else:
        eng.residual(y, eng.ab(st["beta"]), st["z"], st["beta_sq"], None, st["tau"], P, 0, False)
    atz = eng.az(st["z"])
    eng.denoise(st["beta"], atz, st["beta"], s_out, st["tau"], st["beta_sq"], L, M, sq)


FUSED_TC_AMP = True       # tensor-core mode: the whole AMP loop in one C call (lsb_tc_amp_dev_f32)


def _fused_tc_ok(eng, L, M, t_max, F):
    return eng.tc and FUSED_TC_AMP and 2 <= M <= 128 and (L * M) % 32 == 0 and t_max >= 1 and F > 0


def _amp_device(eng, yd, P, L, M, sq, t_max, return_trace=False):
    """t_max AMP iterations on device frames ``yd [F, n]``; returns device tensors
    ``(beta, s[, tau^2 per iteration])``."""
    if _fused_tc_ok(eng, L, M, t_max, yd.shape[0]):
        # three launches per iteration, no Python in the loop: A beta -> residual (+ fp16 split of z) ->
        # A^T z with the denoiser (+ fp16 split of beta) as the GEMM epilogue
        t = eng.torch
        F = yd.shape[0]
        beta, s = eng.empty(F, L * M), eng.empty(F, L * M)
        taus = t.empty((t_max, F), dtype=t.float64, device=eng.dev) if return_trace else None
        _lib.check(_lib.lib().lsb_tc_amp_dev_f32(eng.h, yd.data_ptr(), F, L, M, float(P), float(sq), int(t_max),
                                                 beta.data_ptr(), s.data_ptr(),
                                                 taus.data_ptr() if return_trace else None, None, eng.stream()), "tc_amp")
        if return_trace:
            return beta, s, taus.t().contiguous()
        return beta, s

    def run(yc):
        st = _amp_state(eng, yc, L, M)
        s = eng.empty(yc.shape[0], L * M)
        taus = []
        for t in range(t_max):
            _amp_iteration(eng, yc, st, P, L, M, sq, t == 0, s)
            if return_trace:
                taus.append(st["tau"].clone())
            yield
        if return_trace:
            return st["beta"], s, eng.torch.stack(taus, dim=1)      # [F, t_max]
        return st["beta"], s

    return _drive(eng, yd, run)


def sparc_amp(y, sparc_params, decode_params, A, precision=None, return_trace=False):
    """The normal AMP algorithm (sparc_new.py:885-912).  Returns ``(beta, s)``."""
    P, L, M = sparc_params['P'], sparc_params['L'], sparc_params['M']
    eng = _engine(A, precision)
    single = _is_single(y)
    yd = eng.to_dev(y)
    n = yd.shape[1]
    sq = np.sqrt(n * (P / L))

    out = _amp_device(eng, yd, P, L, M, sq, decode_params['t_max'], return_trace)
    res = (_out(out[0], single), _out(out[1], single))
    if return_trace:
        return res + (_out(out[2], single),)
    return res


def sparc_amp_decide_host(y, sparc_params, decode_params, A, out=None):
    """AMP decode + MAP decision through the host-pointer C-ABI entry ``lsb_tc_amp_host_f32`` (tensor-core mode):
    ``y`` float32 ``[F, n]`` in host memory (numpy, or a pinned CPU torch tensor) -> section decisions ``[F, L]`` int32
    (argmax of the last iteration's s, what sparc_ldpc_decode does without an LDPC code, sparc_new.py:79-80)."""
    import torch
    P, L, M = sparc_params['P'], sparc_params['L'], sparc_params['M']
    eng = _engine(A, "tc")
    n = eng.n
    if isinstance(y, torch.Tensor):
        assert y.device.type == "cpu" and y.dtype == torch.float32 and y.is_contiguous()
        F, y_ptr = y.reshape(-1, n).shape[0], y.data_ptr()
    else:
        y = np.ascontiguousarray(y, dtype=np.float32).reshape(-1, n)
        F, y_ptr = y.shape[0], y.ctypes.data
    if out is None:
        idx = np.empty((F, L), np.int32)
        ptr = idx.ctypes.data
    else:
        idx, ptr = out, out.data_ptr()
    _lib.check(_lib.lib().lsb_tc_amp_host_f32(eng.h, y_ptr, F, L, M, float(P), float(np.sqrt(n * (P / L))),
                                              int(decode_params['t_max']), ptr), "lsb_tc_amp_host_f32")
    return idx


def sparc_amp_single_it(sparc_params, y, A, AT, beta, z, tau_sqr, precision=None):
    """One AMP iteration on host arrays (sparc_new.py:975-990).  Returns ``(beta, z, tau_sqr)``."""
    P, L, M = sparc_params['P'], sparc_params['L'], sparc_params['M']
    eng = _engine(A, precision)
    t = eng.torch
    yd = eng.to_dev(y)
    n = yd.shape[1]
    sq = np.sqrt(n * (P / L))
    st = _amp_state(eng, yd, L, M)
    st["beta"] = eng.to_dev(beta)
    first = np.ndim(z) == 0
    if not first:
        st["z"] = eng.to_dev(z)
        st["tau"] = t.full((1,), float(tau_sqr), dtype=t.float64, device=eng.dev)
        st["beta_sq"] = (st["beta"].double() ** 2).sum(1)
    _amp_iteration(eng, yd, st, P, L, M, sq, first)
    return _out(st["beta"], True), _out(st["z"], True), float(st["tau"][0].item())


def msg_vector_mmse_estimator(s, tau_sqr, n, P_l, M, precision=None):
    """Section-wise MMSE estimator (sparc_new.py:1040-1066), on the GPU."""
    import torch
    _lib.require_device()
    precision = precision or DEFAULT_PRECISION
    dt = torch.float64 if precision == "fp64" else torch.float32
    sd = torch.from_numpy(np.ascontiguousarray(np.asarray(s, dtype=np.float64).reshape(1, -1))).cuda().to(dt)
    L = sd.shape[1] // M
    beta = torch.empty_like(sd)
    tau = torch.full((1,), float(tau_sqr), dtype=torch.float64, device=sd.device)
    bsq = torch.empty(1, dtype=torch.float64, device=sd.device)
    fn = getattr(_lib.lib(), "lsb_amp_denoise_dev_" + ("f64" if precision == "fp64" else "f32"))
    _lib.check(fn(None, sd.data_ptr(), beta.data_ptr(), None, tau.data_ptr(), bsq.data_ptr(), 1, L, M,
                  float(np.sqrt(n * P_l)), torch.cuda.current_stream().cuda_stream), "amp_denoise")
    return _out(beta, True)


# ------------------------------------------------------------------ glue
def beta_estimate_to_bp_probs(beta, L, M, sqrt_nP_l, A=None, precision=None):
    """Section posteriors -> P(bit = 0) per LDPC bit (sparc_new.py:1118-1138), on the GPU."""
    import torch
    _lib.require_device()
    precision = precision or DEFAULT_PRECISION
    dt = torch.float64 if precision == "fp64" else torch.float32
    bd = torch.from_numpy(np.ascontiguousarray(np.asarray(beta, dtype=np.float64).reshape(1, -1))).cuda().to(dt)
    logM = int(np.log2(M))
    probs = torch.empty((1, L * logM), dtype=dt, device=bd.device)
    fn = getattr(_lib.lib(), "lsb_section_to_bit_dev_" + ("f64" if precision == "fp64" else "f32"))
    _lib.check(fn(bd.data_ptr(), probs.data_ptr(), None, L, M, float(sqrt_nP_l), torch.cuda.current_stream().cuda_stream),
               "section_to_bit")
    return _out(probs, True)


def bp_output_to_beta_estimate(ldpc_probs, L, M, sqrt_nP_l, precision=None):
    """P(bit = 0) per bit -> product-form section prior (sparc_new.py:1260-1279), on the GPU
    (through the logit: the kernel takes APP LLRs, sigmoid(logit(p)) = p)."""
    import torch
    _lib.require_device()
    precision = precision or DEFAULT_PRECISION
    dt = torch.float64 if precision == "fp64" else torch.float32
    p = np.asarray(ldpc_probs, dtype=np.float64).reshape(1, -1)
    with np.errstate(divide="ignore"):
        app = np.log(p) - np.log1p(-p)
    ad = torch.from_numpy(np.ascontiguousarray(app)).cuda().to(dt)
    beta = torch.empty((1, L * M), dtype=dt, device=ad.device)
    fn = getattr(_lib.lib(), "lsb_bit_to_section_dev_" + ("f64" if precision == "fp64" else "f32"))
    _lib.check(fn(ad.data_ptr(), None, beta.data_ptr(), None, 1, L, M, float(sqrt_nP_l),
                  torch.cuda.current_stream().cuda_stream), "bit_to_section")
    return _out(beta, True)


def ldpc_bp(ldpc_probs, c, num_its, hard_decision_bool, precision=None):
    """P(bit = 0) -> LLR -> BP per LDPC block -> probabilities or hard info bits
    (sparc_new.py:1162-1193).  Returns ``(ldpc_probs, hard_decision_bits)``."""
    precision = precision or DEFAULT_PRECISION
    eps = 1e-15
    p = np.clip(np.asarray(ldpc_probs, dtype=np.float64), eps, 1 - eps)
    LLR = np.log(p) - np.log(1 - p)
    assert len(LLR) % c.N == 0
    app, _ = c.decode(LLR.reshape(-1, c.N), num_its, precision=precision)
    if hard_decision_bool:
        return 0, np.array((app[:, :c.K].ravel() < 0), dtype=int)
    with np.errstate(over="ignore", invalid="ignore"):
        a = app.ravel()
        return (np.exp(a)) / (1 + np.exp(a)), 0


# ------------------------------------------------------------------ encoder
def sparc_ldpc_encode(sparc_params, ldpc_params, lengths, ldpc_bool, rand_seed):
    """Random user bits -> LDPC codewords -> SPARC codeword (sparc_new.py:15-51).
    Returns ``user_bits, total_bits, beta0, x, A``."""
    P, R, L, M = sparc_params['P'], sparc_params['R'], sparc_params['L'], sparc_params['M']
    logM = int(np.log2(M))
    if ldpc_bool:
        L_unprotected, k_ldpc, mults = lengths['L_unprotected'], lengths['k_ldpc'], lengths['mults']
        unprotected_bit_len = int(L_unprotected * logM)
        user_bits_len = int(k_ldpc + unprotected_bit_len)
    else:
        user_bits_len = L * logM
    rng = np.random.default_rng(rand_seed)
    user_bits = rng.integers(0, 2, size=user_bits_len)
    if ldpc_bool:
        total_bits = encode_ldpc(user_bits, ldpc_params, mults, unprotected_bit_len)
    else:
        total_bits = user_bits.astype(bool)
    encoded_bit_len = total_bits.size
    assert encoded_bit_len == L * logM
    n = int(encoded_bit_len / R)
    P_l = P / L
    beta0 = bin_arr_2_msg_vector(total_bits, M, n, P_l)
    A = create_design_matrix(L, M, n, rand_seed)
    eng = _engine(A, "fp64")
    x = _out(eng.ab(eng.to_dev(beta0)), True)
    return user_bits, total_bits, beta0, x, A


# ------------------------------------------------------------------ decoders
def _finish_bits(bits, single):
    b = bits.cpu().numpy().astype(int)
    return b[0] if single else b


def sparc_ldpc_decode(y, sparc_params, ldpc_params, decode_params, ldpc_bool, lengths, A, precision=None):
    """AMP fully, then BP fully (sparc_new.py:53-82).  Returns the decoded user bits."""
    P, R, L, M = sparc_params['P'], sparc_params['R'], sparc_params['L'], sparc_params['M']
    eng = _engine(A, precision)
    single = _is_single(y)
    yd = eng.to_dev(y)
    n = yd.shape[1]
    sq = np.sqrt(n * (P / L))
    if ldpc_bool and lengths['L_unprotected'] != 0:
        raise NotImplementedError("L_unprotected > 0 does not work in the reference either "
                                  "(sparc_new.py:68-73 reshapes with the full L)")
    c = code(ldpc_params["standard"], ldpc_params["rate"], ldpc_params["z"]) if ldpc_bool else None
    logM = int(round(np.log2(M)))

    def finish(beta, s, Fc):
        if ldpc_bool:
            _, llr = eng.section_to_bit(beta, L, M, sq, want_probs=False)
            return eng.hard_bits(eng.bp(llr, c, 200), c)
        idx = s.reshape(Fc, L, M).argmax(dim=2)                   # MAP on s (sparc_new.py:79-80)
        shifts = eng.torch.arange(logM - 1, -1, -1, device=eng.dev)
        return ((idx.unsqueeze(-1) >> shifts) & 1).reshape(Fc, -1).to(eng.torch.uint8)

    def run(yc):
        st = _amp_state(eng, yc, L, M)
        s = eng.empty(yc.shape[0], L * M)
        for t in range(decode_params['t_max']):
            _amp_iteration(eng, yc, st, P, L, M, sq, t == 0, s)
            yield
        return finish(st["beta"], s, yc.shape[0])

    if _fused_tc_ok(eng, L, M, decode_params['t_max'], yd.shape[0]):
        beta, s = _amp_device(eng, yd, P, L, M, sq, decode_params['t_max'])     # the AMP loop as one C call
        bits = finish(beta, s, yd.shape[0])
    else:
        bits = _drive(eng, yd, run)
    if ldpc_bool:
        return _finish_bits(bits, single)
    bits = bits.bool().cpu().numpy()
    return bits[0] if single else bits


def naively_integrated_decoder(y, sparc_params, ldpc_params, decode_params, A, precision=None, return_trace=False):
    """One AMP iteration followed by 6 BP iterations, t_max times; finished by 200 BP
    iterations (sparc_new.py:257-282).  Returns the hard-decision user bits."""
    P, R, L, M = sparc_params['P'], sparc_params['R'], sparc_params['L'], sparc_params['M']
    eng = _engine(A, precision)
    single = _is_single(y)
    yd = eng.to_dev(y)
    n = yd.shape[1]
    sq = np.sqrt(n * (P / L))
    c = code(ldpc_params["standard"], ldpc_params["rate"], ldpc_params["z"])
    t_max = decode_params['t_max']

    def run(yc):
        st = _amp_state(eng, yc, L, M)
        taus = []
        for i in range(t_max):
            _amp_iteration(eng, yc, st, P, L, M, sq, i == 0)
            if return_trace:
                taus.append(st["tau"].clone())
            _, llr = eng.section_to_bit(st["beta"], L, M, sq, want_probs=False)
            if i != t_max - 1:
                app = eng.bp(llr, c, 6)
                _, st["beta"] = eng.bit_to_section(app, L, M, sq, beta_sq=st["beta_sq"], want_probs=False)
            else:
                bits = eng.hard_bits(eng.bp(llr, c, 200), c)
            yield
        if return_trace:
            return bits, eng.torch.stack(taus, dim=1)
        return bits

    res = _drive(eng, yd, run)
    if return_trace:
        return _finish_bits(res[0], single), _out(res[1], single)
    return _finish_bits(res, single)


def integrated_decoder(y, sparc_params, ldpc_params, decode_params, A, precision=None, return_trace=False):
    """AMP with the BP-aware denoiser and its corrected Onsager term
    (sparc_new.py:472-502, 709-735, 824-883).  Returns the hard-decision user bits."""
    P, R, L, M = sparc_params['P'], sparc_params['R'], sparc_params['L'], sparc_params['M']
    eng = _engine(A, precision)
    t = eng.torch
    single = _is_single(y)
    yd = eng.to_dev(y)
    F, n = yd.shape
    sq = np.sqrt(n * (P / L))
    c = code(ldpc_params["standard"], ldpc_params["rate"], ldpc_params["z"])
    t_max = decode_params['t_max']
    num_its, num_its_final = 6, 200

    def run(yc):
        Fc = yc.shape[0]
        beta = t.zeros((Fc, L * M), dtype=eng.dt, device=eng.dev)
        z = eng.empty(Fc, n)
        tau = t.ones(Fc, dtype=t.float64, device=eng.dev)
        dummy_sq = t.zeros(Fc, dtype=t.float64, device=eng.dev)
        alpha_w = eng.empty(Fc, L * M)
        taus, sums = [], []
        vk = vk0 = None
        for it in range(t_max):
            if it != 0:
                ons = eng.onsager(beta, vk, vk0, alpha_w, tau, L, M, sq)           # uses tau of the previous iteration
                if return_trace:
                    sums.append(ons.clone())
                eng.residual(yc, eng.ab(beta), z, None, ons, tau, P, 1, False)
            else:
                eng.residual(yc, None, z, None, None, tau, P, 1, True)
            atz = eng.az(z)
            eng.denoise(beta, atz, alpha_w, None, tau, dummy_sq, L, M, sq)         # eta step one: weighted alpha
            if return_trace:
                taus.append(tau.clone())
            vk0, llr = eng.section_to_bit(alpha_w, L, M, sq)                       # step two
            if it != t_max - 1:
                app = eng.bp(llr, c, num_its)                                      # step three
                vk, beta = eng.bit_to_section(app, L, M, sq)                       # step four
            else:
                bits = eng.hard_bits(eng.bp(llr, c, num_its_final), c)
            yield
        if return_trace:
            ss = t.stack(sums, dim=1) if sums else t.zeros((Fc, 0), dtype=t.float64, device=eng.dev)
            return bits, t.stack(taus, dim=1), ss
        return bits

    res = _drive(eng, yd, run)
    if return_trace:
        return _finish_bits(res[0], single), _out(res[1], single), _out(res[2], single)
    return _finish_bits(res, single)

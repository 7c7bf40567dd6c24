"""State evolution of SPARC AMP decoding on the GPU: drop-in for sparc_public/sparc_se.py
(``check_code_params``, ``sparc_se_E``, ``sparc_se`` -- same arguments and returns).

The recursion over iterations is a handful of scalars per step and stays on the host;
the Monte-Carlo expectation over ``mc_samples x M`` Gaussians (all of the reference's
run time) is one kernel launch per iteration covering every column block
(csrc/sparc_se.cu), for every SPARC type of the reference (real or complex design,
K-PSK with K up to 64).  No CPU fallback.

As in the reference (sparc_se.py:150-153) the samples come from NumPy's global generator
(``np.random.randn``) and are drawn once per call, so ``np.random.seed(s)`` before the call
reproduces the reference's numbers.
"""
from copy import copy

import numpy as np

from . import _lib
from .sparc import create_base_matrix, is_power_of_2

_FLAGS = ('complex', 'modulated', 'power_allocated', 'spatially_coupled')
_FLOAT = (float, np.float64)


def check_code_params(code_params):
    """Validates and normalises ``code_params`` IN PLACE for state evolution: like
    sparc.check_code_params but without L and the divisibility rules (sparc_se.py:14-80)."""
    kept = {}
    for flag in _FLAGS:
        value = code_params.setdefault(flag, False)
        assert type(value) == bool, "'{}' must be boolean".format(flag)
        kept[flag] = copy(value)

    def need(keys):
        if any(k not in code_params for k in keys):
            raise Exception('Need code parameters {}.'.format(list(keys)))
        for k in keys:
            kept[k] = copy(code_params[k])

    need(['P', 'R', 'M'])
    P, R, M = code_params['P'], code_params['R'], code_params['M']
    assert type(P) in _FLOAT and P > 0
    assert type(R) in _FLOAT and R > 0
    assert type(M) == int and M > 0 and is_power_of_2(M)
    if kept['modulated']:
        need(['K'])
        K = code_params['K']
        assert type(K) == int and K > 1 and is_power_of_2(K)
        if not kept['complex']:
            assert K == 2, 'Real-modulated SPARCs requires K=2'
    if kept['power_allocated']:
        need(['B', 'R_PA_ratio'])
        B, ratio = code_params['B'], code_params['R_PA_ratio']
        assert type(B) == int and B > 1
        assert type(ratio) in _FLOAT
        assert ratio >= 0
    if kept['spatially_coupled']:
        need(['omega', 'Lambda'])
        omega, Lambda = code_params['omega'], code_params['Lambda']
        assert type(omega) == int and omega > 1
        assert type(Lambda) == int and Lambda >= (2 * omega - 1)
    code_params.clear()
    code_params.update(kept)


class _Samples:
    """The Monte-Carlo draws of one sparc_se call, resident on the device."""

    def __init__(self, u, precision="fp64"):
        import torch
        _lib.require_device()
        u = np.ascontiguousarray(u)
        self.cplx = np.iscomplexobj(u)
        self.mc, self.M = u.shape
        self.f32 = precision == "fp32"
        if self.cplx:
            dt = torch.complex64 if self.f32 else torch.complex128
        else:
            dt = torch.float32 if self.f32 else torch.float64
        self.dev = torch.from_numpy(u).to("cuda").to(dt).contiguous()

    def expect(self, taus, K):
        import torch
        taus = np.ascontiguousarray(np.atleast_1d(np.asarray(taus, dtype=np.float64)))
        E = np.empty_like(taus)
        if (K > 2) != self.cplx:
            raise ValueError("K > 2 needs complex samples, K <= 2 real ones (sparc_se.py:150-153)")
        name = {(False, False): "f64", (False, True): "f32", (True, False): "c128", (True, True): "c64"}[(self.cplx, self.f32)]
        fn = getattr(_lib.lib(), "lsb_sparc_se_expect_" + name)
        with torch.cuda.device(self.dev.device):
            _lib.check(fn(self.dev.data_ptr(), self.mc, self.M, int(K), taus.ctypes.data_as(_lib.c_f64p), taus.size,
                          E.ctypes.data_as(_lib.c_f64p), torch.cuda.current_stream().cuda_stream), "lsb_sparc_se_expect")
        return E


def sparc_se_E(tau, K, u, precision="fp64"):
    """Monte-Carlo estimate of E[beta_1-hat] at noise level ``tau`` (sparc_se.py:82-118)."""
    s = u if isinstance(u, _Samples) else _Samples(u, precision)
    return float(s.expect([tau], K)[0])


def sparc_se(awgn_var, code_params, t_max, mc_samples, precision="fp64"):
    """State evolution for a general base matrix (sparc_se.py:120-183).
    Returns ``psi`` (``[t_max]`` or ``[t_max, Lc]``) and the final ``tau``."""
    check_code_params(code_params)
    tmp = code_params.copy()
    tmp.update({'awgn_var': awgn_var})
    W = create_base_matrix(**tmp)
    assert 0 <= W.ndim <= 2
    P, R, M = map(code_params.get, ['P', 'R', 'M'])
    K = code_params['K'] if code_params['modulated'] else 1
    if K > 64:
        raise NotImplementedError("K-PSK with K > 64 is not built")
    if code_params['complex']:
        R /= 2                    # complex SPARCs: rate per dimension (sparc_se.py:141-142)

    if W.ndim == 0:
        psi = np.ones(t_max)
    else:
        Lr, Lc = (1, W.size) if W.ndim == 1 else W.shape
        psi = np.ones((t_max, Lc))
    if K > 2:                     # must be complex (:150-153); two draws, real parts first
        u = np.random.randn(mc_samples, M) + 1j * np.random.randn(mc_samples, M)
    else:
        u = np.random.randn(mc_samples, M)
    samples = _Samples(u, precision)

    for t in range(t_max - 1):
        if t > 0:
            tau_prev = np.copy(tau)
        if W.ndim == 0:
            tau = (np.log(2) * R / np.log(K * M)) * (awgn_var / P + psi[t])
        else:
            phi = awgn_var + np.dot(W, psi[t]) / Lc
            tau = (np.log(2) * R * Lr / np.log(K * M)) / np.dot(W.T, 1 / phi)
        if (t > 0) and np.allclose(tau, tau_prev, rtol=1e-6, atol=0):
            psi[t:] = psi[t]
            break
        E = samples.expect(tau, K)
        psi[t + 1] = 1 - (E[0] if W.ndim == 0 else E)
    return psi, tau

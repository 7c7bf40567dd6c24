"""CPU restatement of the reference's state evolution (sparc_public/sparc_se.py) for
every K.  ORACLE / TEST INFRASTRUCTURE ONLY: imported from tests/ only,
never from ldpc_sparc_b200/.  Pinned against the unmodified reference by
oracle/make_golden_se.py -> tests/golden/se_reference.npz (tests/test_oracle_se.py).
"""
import numpy as np

from sparc_oracle import create_base_matrix


def se_expectation(tau, K, u):
    """sparc_se.py:82-118 (sparc_se_E), K = 1 / 2."""
    itau = 1.0 / tau
    rtau = np.sqrt(itau)
    if K == 1:
        a = np.exp(itau + rtau * u[:, 0])
        c = np.exp(rtau * u[:, 1:])
    elif K == 2:
        a = np.sinh(itau + rtau * u[:, 0])
        c = np.cosh(rtau * u[:, 1:])
    elif K == 4:
        a = np.sinh(itau + rtau * u[:, 0].real)
        b = np.cosh(itau + rtau * u[:, 0].real) + np.cosh(rtau * u[:, 0].imag)
        c = np.cosh(rtau * u[:, 1:].real) + np.cosh(rtau * u[:, 1:].imag)
        return (a / (b + c.sum(axis=1))).mean()
    else:
        from sparc_oracle import psk_constel
        con = psk_constel(K)
        tb = np.exp(np.real((itau + rtau * u[:, 0])[None, :] * con.conj()[:, None]))
        ta = con.real[:, None] * tb
        tc = np.exp(np.real((rtau * u[:, 1:])[None, :, :] * con.conj()[:, None, None]))
        return (ta.mean(axis=0) / (tb.mean(axis=0) + tc.mean(axis=0).sum(axis=1))).mean()
    return (a / (a + c.sum(axis=1))).mean()


def state_evolution(awgn_var, code_params, t_max, u):
    """sparc_se.py:120-183 with the Monte-Carlo samples ``u [mc, M]`` passed in."""
    cp = dict(code_params)
    W = create_base_matrix(**dict(cp, awgn_var=awgn_var))
    P, R, M = cp['P'], cp['R'], cp['M']
    K = cp['K'] if cp.get('modulated') else 1
    if cp.get('complex'):
        R = R / 2
    if W.ndim == 0:
        psi = np.ones(t_max)
    else:
        Lr, Lc = (1, W.size) if W.ndim == 1 else W.shape
        psi = np.ones((t_max, Lc))
    tau = None
    for t in range(t_max - 1):
        tau_prev = tau
        if W.ndim == 0:
            tau = (np.log(2) * R / np.log(K * M)) * (awgn_var / P + psi[t])
        else:
            phi = awgn_var + np.dot(W, psi[t]) / Lc
            tau = (np.log(2) * R * Lr / np.log(K * M)) / np.dot(W.T, 1 / phi)
        if t > 0 and np.allclose(tau, tau_prev, rtol=1e-6, atol=0):
            psi[t:] = psi[t]
            break
        if W.ndim == 0:
            psi[t + 1] = 1 - se_expectation(tau, K, u)
        else:
            psi[t + 1] = [1 - se_expectation(tc, K, u) for tc in tau]
    return psi, tau

"""Parameter bookkeeping of the LDPC-protected SPARC experiments (host only; mirrors
ldpc_sparc/param_calc.py:11-60 -- same names, arguments and return tuples)."""
import numpy as np

from .ldpc import code


def param_calc(mults, logM, standard, ldpc_rate, int_rate, z, R_sparc_ldpc):
    """Fully protected scheme: ``mults`` LDPC blocks fill the SPARC exactly.
    Returns ``overall_rate, L_sparc, L_sparc_ldpc, lengths``."""
    k = code(standard, ldpc_rate, z).K * mults
    coded = k / int_rate
    assert coded % 1 == 0, "ldpc_bits_len must be an integer"
    coded = int(coded)
    assert k % logM == 0
    assert coded % logM == 0
    n = int(coded / R_sparc_ldpc)
    lengths = {'k_ldpc': k, 'mults': mults, 'L_unprotected': 0}
    return k / n, int(k / logM), int(coded / logM), lengths


def param_calc_semi_protected(R, mults, percent_protected, M, standard, ldpc_rate, int_rate, z):
    """Part of the sections carry unprotected bits.
    Returns ``L_sparc, R_sparc_ldpc, L_sparc_ldpc, lengths, updated_rate``."""
    logM = np.log2(M)
    k_ldpc = code(standard, ldpc_rate, z).K * mults
    n_ldpc = int(k_ldpc / int_rate)
    assert n_ldpc % logM == 0
    unprotected_bits = int((k_ldpc * (1 - percent_protected)) / percent_protected)
    unprotected_bits = np.ceil(unprotected_bits / logM) * logM
    coded = n_ldpc + unprotected_bits
    L_sparc_ldpc = int(coded / logM)
    L_unprotected = int(unprotected_bits / logM)
    assert coded >= n_ldpc
    k = k_ldpc + unprotected_bits
    n = int(k / R)
    lengths = {'k_ldpc': k_ldpc, 'mults': mults, 'L_unprotected': L_unprotected}
    return int(k // logM), coded / n, L_sparc_ldpc, lengths, k / n

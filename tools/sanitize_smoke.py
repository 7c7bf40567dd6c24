#!/usr/bin/env python
"""Small invocations of every kernel family (a quick all-kernels smoke; also the workload to
put under a memory checker where one is available):
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
Sizes are tiny (the tools slow kernels down 10-100x); correctness is the test suite's job."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from ldpc_sparc_b200 import ldpc, sparc, sparc_new, sparc_se
    which = sys.argv[1:] or ["ldpc", "sparc", "complex", "dense", "se"]
    rng = np.random.RandomState(0)
    if "ldpc" in which:
        c = ldpc.code("802.11n", "1/2", 27)
        x = c.encode(rng.randint(0, 2, (8, c.K)))
        llr = 2.0 * ((1 - 2.0 * x) + 0.8 * rng.randn(8, c.N)) / 0.64
        for prec in ("fp32", "fp64"):
            app, it = c.decode(llr, 20, precision=prec)
        c.decode(llr[0], 20, "sumprod")
        c.decode(llr[:2], 20, "minsum")
        xd = c.encode_device(torch.from_numpy(rng.randint(0, 2, (8, c.K))).cuda())
        print("ldpc ok", it[:4], int(xd.sum()))
    if "sparc" in which:
        for cp in (dict(P=15.0, R=1.3, L=32, M=16), dict(P=15.0, R=1.4, L=64, M=32, modulated=True, K=2, spatially_coupled=True,
                                                         omega=2, Lambda=4, power_allocated=True, B=2, R_PA_ratio=0.5)):
            bits, beta0, x, Ab, Az = sparc.sparc_encode(cp, 1.0, [1, 2])
            y = np.stack([x + rng.randn(x.size), x + rng.randn(x.size)])
            for prec in ("fp32", "fp64"):
                out = sparc.sparc_decode(y, cp, dict(t_max=8), 1.0, [1, 2], beta0, Ab, Az, precision=prec)
            Az(y[0])
        print("sparc ok", out[2])
    if "complex" in which:
        cp = dict(P=15.0, R=2.0, L=32, M=16, complex=True, modulated=True, K=8, spatially_coupled=True, omega=2, Lambda=4)
        bits, beta0, x, Ab, Az = sparc.sparc_encode(cp, 1.0, [1, 2])
        y = np.stack([x + 0.7 * (rng.randn(x.size) + 1j * rng.randn(x.size))] * 2)
        for prec in ("fp32", "fp64"):
            out = sparc.sparc_decode(y, cp, dict(t_max=8, phi_est_method=2), 1.0, [1, 2], beta0, Ab, Az, precision=prec)
        Az(y[0])
        print("complex ok", out[2])
    if "dense" in which:
        from ldpc_sparc_b200.param_calc import param_calc
        rate, L_sparc, L_sl, lengths = param_calc(1, 2, '802.11n', '1/2', 1 / 2, 27, 1)
        spl = {'P': 19.44, 'R': 1, 'L': L_sl, 'M': 4}
        lp = {'standard': '802.11n', 'rate': '1/2', 'z': 27, 'int_rate': 1 / 2, 'mults': 1}
        ub, tb, b0, x, A = sparc_new.sparc_ldpc_encode(spl, lp, lengths, True, 3)
        y = np.stack([x + 1.5 * rng.randn(x.size)] * 3)
        for prec in ("tc", "fp32", "fp64"):
            sparc_new.sparc_ldpc_decode(y, spl, lp, {'t_max': 4}, True, lengths, A, precision=prec)
            sparc_new.naively_integrated_decoder(y, spl, lp, {'t_max': 3}, A, precision=prec)
            bits = sparc_new.integrated_decoder(y, spl, lp, {'t_max': 3}, A, precision=prec)
        print("dense ok", bits.shape)
    if "se" in which:
        np.random.seed(0)
        psi, _ = sparc_se.sparc_se(1.0, dict(P=15.0, R=1.3, M=32), 6, 700)
        psi2, _ = sparc_se.sparc_se(1.0, dict(P=15.0, R=2.6, M=16, complex=True, modulated=True, K=8), 6, 300)
        print("se ok", psi[-1], psi2[-1])
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()

"""Error statistics of BASELINE config 1 (sparc_public PA SPARC, L=1024 M=32 R=1.5, awgn_var 1,
t_max 25) from the numpy oracle (bit-identical to the reference, tests/test_oracle_sparc.py):
section / frame error rates and the mean iteration count over N random frames, for the
statistical parity test of the batched GPU path at full size (tests/test_sims_gpu.py).

    OMP_NUM_THREADS=1 python oracle/make_golden_stats.py [frames]   ->  tests/golden/c1_statistics.json
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import sparc_oracle as so  # noqa: E402

C1 = dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9)
_st = {}


def _init():
    cp = so.normalise_code_params(C1)
    _, _, _, design, n = so.sparc_encode(cp, 1.0, [1, 2])
    _st.update(cp=dict(cp, n=n), design=design)


def _frame(seed):
    rng = np.random.RandomState(1000 + seed)
    L, M = C1["L"], C1["M"]
    idx = rng.randint(0, M, L)
    beta0 = np.zeros(L * M)
    beta0[np.arange(L) * M + idx] = 1
    y = _st["design"].Ab(beta0) + rng.randn(_st["design"].n)
    out = so.sparc_decode(y, _st["cp"], dict(t_max=25), 1.0, _st["design"], beta0)
    return int((out[1].reshape(L, M).argmax(1) != idx).sum()), int(out[2])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    with mp.get_context("fork").Pool(os.cpu_count() or 1, initializer=_init) as pool:
        res = pool.map(_frame, range(n), chunksize=4)
    errs, T = np.array([r[0] for r in res]), np.array([r[1] for r in res])
    out = dict(code_params=C1, awgn_var=1.0, t_max=25, design_seed=[1, 2], frames=n,
               section_errors=int(errs.sum()), frames_in_error=int((errs > 0).sum()), mean_t_final=float(T.mean()),
               std_t_final=float(T.std()), section_errors_per_frame_var=float(errs.var()))
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "c1_statistics.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

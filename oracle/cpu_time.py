#!/usr/bin/env python
"""Times the CPU oracle (numpy restatement of the reference) on a bounded sample of a
BASELINE config, for bench.py's cpu_baseline legs.  ORACLE / TEST INFRASTRUCTURE ONLY: run
by bench.py as a subprocess; nothing under ldpc_sparc_b200/ imports it.

    python oracle/cpu_time.py c1 --frames 32 --procs 16      # sparc_public AMP decode, config 1
    python oracle/cpu_time.py c3 --frames 2                   # dense-design AMP, config 3
    python oracle/cpu_time.py c4                              # SC wave decode, config 4 (one frame, ~20 s)
    python oracle/cpu_time.py c5                              # LDPC-SPARC integrated decoder, config 5 (one frame)

Prints one JSON object: decoded info bits/s over the sample, decode calls only (encoding and
design construction are outside the timed region, as SURVEY 8(d) asks)."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

C1 = dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9)
_state = {}


def _c1_init():
    import sparc_oracle as so
    cp = so.normalise_code_params(C1)
    bits, beta0, x, design, n = so.sparc_encode(cp, 1.0, [1, 2])
    _state.update(so=so, cp=dict(cp, n=n), design=design, x=x)


def _c1_decode(seed):
    so = _state["so"]
    y = _state["x"] + np.random.RandomState(seed).randn(_state["x"].size)
    t0 = time.perf_counter()
    out = so.sparc_decode(y, _state["cp"], dict(t_max=25), 1.0, _state["design"])
    return time.perf_counter() - t0, int(out[2])


def time_c1(frames, procs):
    with mp.get_context("fork").Pool(procs, initializer=_c1_init) as pool:
        pool.map(_c1_decode, range(procs))                      # warm every worker
        t0 = time.perf_counter()
        res = pool.map(_c1_decode, range(100, 100 + frames), chunksize=1)
        wall = time.perf_counter() - t0
    bits = frames * C1["L"] * 5
    return {"value": bits / wall, "unit": "info bits/s", "cores": procs, "kind": "port",
            "sample": "%d frames of config 1 (L=1024 M=32 R=1.5 PA, t_max 25, mean t_final %.1f), numpy oracle of "
                      "sparc_decode, %d processes x 1 BLAS thread, %.1f s (%.2f s per frame per core)"
                      % (frames, float(np.mean([r[1] for r in res])), procs, wall, float(np.mean([r[0] for r in res])))}


def time_c3(frames):
    import sophie_oracle as sph
    L, M, n, P = 512, 64, 3072, 19.44
    rng = np.random.default_rng(0)
    A = rng.normal(0, 1 / np.sqrt(n), (n, L * M))
    sq = np.sqrt(n * P / L)
    total = 0.0
    for f in range(frames):
        beta0 = np.zeros(L * M)
        beta0[np.arange(L) * M + rng.integers(0, M, L)] = sq
        y = A @ beta0 + np.sqrt(P / 6.0) * rng.standard_normal(n)
        t0 = time.perf_counter()
        sph.sparc_amp(y, dict(P=P, L=L, M=M), dict(t_max=25), A)
        total += time.perf_counter() - t0
    return {"value": frames * L * 6 / total, "unit": "info bits/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d frames of config 3 (A 3072x32768 float64, t_max 25), numpy oracle of sparc_new.sparc_amp, "
                      "BLAS on all host threads, %.1f s" % (frames, total)}


def time_c4(frames):
    import sparc_oracle as so
    cp = so.normalise_code_params(dict(P=15.0, R=1.5, L=2048, M=512, spatially_coupled=True, omega=6, Lambda=32))
    bits, beta0, x, design, n = so.sparc_encode(cp, 1.0, [1, 2])
    total, its = 0.0, []
    for f in range(frames):
        y = x + np.random.RandomState(50 + f).randn(x.size)
        t0 = time.perf_counter()
        out = so.sparc_decode(y, dict(cp, n=n), dict(t_max=40), 1.0, design)
        total += time.perf_counter() - t0
        its.append(int(out[2]))
    return {"value": frames * 2048 * 9 / total, "unit": "info bits/s", "cores": 1, "kind": "port",
            "sample": "%d frame(s) of config 4 (SC wave, L=2048 M=512, w=65536, t_max 40, t_final %s), numpy oracle of "
                      "sparc_decode, one process, %.1f s" % (frames, its, total)}


def time_c5(frames):
    import ldpc_oracle as orc
    import sophie_oracle as sph
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ldpc_sparc_b200 import ldpc as facade          # graph construction is host code as well
    c = facade.code("802.16", "1/2", 150)
    oc = sph.Code(c)
    L = n = 14400
    M, P, mults = 2, 19.44, 4
    rng = np.random.default_rng(0)
    A = rng.normal(0, 1 / np.sqrt(n), (n, L * M))
    sq = np.sqrt(n * P / L)
    total = 0.0
    for f in range(frames):
        user = rng.integers(0, 2, mults * c.K)
        coded = np.concatenate([c.encode(u) for u in user.reshape(mults, c.K)])
        beta0 = np.zeros(L * M)
        beta0[np.arange(L) * M + coded] = sq
        y = A @ beta0 + np.sqrt(P / 4.0) * rng.standard_normal(n)
        t0 = time.perf_counter()
        sph.integrated_decoder(y, dict(P=P, R=1, L=L, M=M), oc, dict(t_max=25), A)
        total += time.perf_counter() - t0
    return {"value": frames * mults * c.K / total, "unit": "user bits/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "%d frame(s) of config 5 (802.16 r1/2 z=150 x 4, M=2, L=n=14400, t_max 25, snr 4), numpy oracle of "
                      "integrated_decoder, BLAS on all host threads, %.1f s" % (frames, total)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    if args.config == "c1":
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        print(json.dumps(time_c1(args.frames or 2 * args.procs, args.procs)))
    elif args.config == "c3":
        print(json.dumps(time_c3(args.frames or 2)))
    elif args.config == "c4":
        print(json.dumps(time_c4(args.frames or 1)))
    else:
        print(json.dumps(time_c5(args.frames or 1)))


if __name__ == "__main__":
    main()

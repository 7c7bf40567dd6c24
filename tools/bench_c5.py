#!/usr/bin/env python
"""Throughput of the LDPC-protected dense-design SPARC decoders at BASELINE config 5
(performance_plots_general.py:21-47): 802.16 r1/2 z=150, mults=4, M=2 -> L = n = 14400,
A 14400 x 28800, t_max = 25; the four schemes of the reference's sweep, frames generated,
decoded and counted on the GPU (sparc_sim_new.performance_sweep).

Prints one JSON object with frames/s and decoded user bits/s per scheme.  --small uses the
parity variant (mults=1, L = n = 3600)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256, help="frames per SNR point and scheme")
    ap.add_argument("--chunk", type=int, default=256)
    ap.add_argument("--snrs", type=float, nargs="+", default=[2.0, 6.0])
    ap.add_argument("--precision", default="tc")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--sims", nargs="+", default=None)
    args = ap.parse_args()
    import torch
    from ldpc_sparc_b200 import sparc_sim_new
    from ldpc_sparc_b200.param_calc import param_calc

    mults = 1 if args.small else 4
    overall, L_sparc, L_sl, lengths = param_calc(mults, 1, '802.16', '1/2', 1 / 2, 150, 1)
    sp = {'P': 19.44, 'R': overall, 'L': L_sparc, 'M': 2}
    spl = {'P': 19.44, 'R': 1, 'L': L_sl, 'M': 2}
    lp = {'standard': '802.16', 'rate': '1/2', 'z': 150, 'int_rate': 1 / 2, 'mults': mults}
    sims = tuple(args.sims) if args.sims else sparc_sim_new.SIMS
    res = {"workload": "config 5: 802.16 r1/2 z=150 x %d blocks, M=2, L=n=%d, t_max=25, %d frames per point, SNR %s"
                       % (mults, L_sl, args.frames, args.snrs), "precision": args.precision, "schemes": {}}
    for name in sims:
        kw = dict(sims=(name,), chunk=args.chunk, precision=args.precision)
        # warm-up = the same sweep (design upload, scratch / staging growth, allocator pools), then best of two
        sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': 25}, args.snrs, args.frames, seed=1, **kw)
        dt = float("inf")
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': 25}, args.snrs, args.frames, seed=1, **kw)
            torch.cuda.synchronize()
            dt = min(dt, time.perf_counter() - t0)
        n_frames = args.frames * len(args.snrs)
        res["schemes"][name] = {"frames_per_s": n_frames / dt, "user_bits_per_s": float(out["bits"].sum()) / dt,
                                "ber": [float(v) for v in out["ber_store_averages"][0]], "seconds": dt}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Throughput of the LDPC-protected dense-design SPARC decoders at BASELINE config 5
(performance_plots_general.py:21-47): 802.16 r1/2 z=150, mults=4, M=2 -> L = n = 14400, A 14400 x 28800, t_max = 25;
frames generated, decoded and counted on the GPU (sparc_sim_new.performance_sweep).  Prints one JSON object
(tools/workloads.c5_workload) with frames/s, decoded user bits/s and the tensor roofline.  --small uses the parity
variant (mults=1, L = n = 3600)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256, help="frames per SNR point and scheme")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--snrs", type=float, nargs="+", default=[2.0, 6.0])
    ap.add_argument("--precision", default="tc")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--sims", nargs="+", default=["SPARC+LDPC", "Integrated_decoder"])
    args = ap.parse_args()
    ctx = workloads.single_rank_ctx()
    print(json.dumps(workloads.c5_workload(ctx, args.frames, args.steps, args.warmup, args.precision, args.small,
                                           tuple(args.sims), tuple(args.snrs))))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Throughput of the batched SPARC AMP decoder on BASELINE config 1 (sparc_public, power-allocated, L=1024 M=32
R=1.5, sub-sampled DCT design, t_max=25, awgn_var=1) or config 4 (SC wave decode, --config c4).

Prints one JSON object (tools/workloads.amp_workload): decoded info bits/s with inputs resident, the same through the
host-pointer C-ABI entry (e2e), frames/s, mean iterations, section error rate, SM clocks during the timed region and
the achieved fraction of the HBM state-streaming model of SURVEY 8(d).
`python tools/bench_amp.py [--frames F] [--steps K] [--warmup W] [--config c1|c4] [--atol A] [--no-e2e]`."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--atol", type=float, default=None, help="early-stop atol (default: the reference's 2e-15)")
    ap.add_argument("--config", default="c1", choices=["c1", "c4"],
                    help="c1: PA L=1024 M=32 R=1.5; c4: SC wave L=2048 M=512 R=1.5 omega=6 Lambda=32 t_max=40")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    ctx = workloads.single_rank_ctx()
    print(json.dumps(workloads.amp_workload(ctx, args.config, args.frames, args.steps, args.warmup, args.precision,
                                            args.atol, e2e=not args.no_e2e)))


if __name__ == "__main__":
    main()

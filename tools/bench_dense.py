#!/usr/bin/env python
"""Throughput of the dense-Gaussian-design AMP decoder (BASELINE config 3: sparc_sophie design, L=512 M=64, R=1 ->
n=3072, A = 3072 x 32768, P=19.44, t_max=25), F frames sharing one design.  Prints one JSON object
(tools/workloads.dense_workload): decoded info bits/s and the tensor roofline of SURVEY 8(d): 402.7 MFLOP per
frame-iteration against the measured sustained bf16 peak.  precision: tc (tcgen05 split-fp16) | fp32 (SIMT)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="tc")
    ap.add_argument("--L", type=int, default=512)
    ap.add_argument("--M", type=int, default=64)
    ap.add_argument("--snr", type=float, default=6.0)
    args = ap.parse_args()
    ctx = workloads.single_rank_ctx()
    print(json.dumps(workloads.dense_workload(ctx, args.frames, args.steps, args.warmup, args.precision, args.L, args.M, args.snr)))


if __name__ == "__main__":
    main()

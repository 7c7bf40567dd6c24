#!/usr/bin/env python
"""Throughput of the dense-Gaussian-design AMP decoder (BASELINE config 3: sparc_sophie
design, L=512 M=64, R=1 -> n=3072, A = 3072 x 32768, P=19.44, t_max=25), F frames sharing
one design.  Reports decoded info bits/s and the tensor roofline of SURVEY 8(d):
402.7 MFLOP per frame-iteration (A beta and A^T z; the first iteration has no A beta)
against the measured bf16 peak.  precision: tc (tcgen05 split-bf16) | fp32 (SIMT)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--precision", default="tc")
    ap.add_argument("--L", type=int, default=512)
    ap.add_argument("--M", type=int, default=64)
    ap.add_argument("--snr", type=float, default=6.0)
    args = ap.parse_args()
    import torch
    from ldpc_sparc_b200 import sparc_new

    L, M, P, t_max = args.L, args.M, 19.44, 25
    logM = int(np.log2(M))
    n = L * logM
    sp = dict(P=P, R=1.0, L=L, M=M)
    dp = dict(t_max=t_max)
    A = sparc_new.create_design_matrix(L, M, n, 0)
    eng = sparc_new._engine(A, args.precision)
    F = args.frames
    gen = torch.Generator(device="cuda").manual_seed(0)
    idx0 = torch.randint(0, M, (F, L), device="cuda", generator=gen)
    sq = float(np.sqrt(n * P / L))
    beta0 = torch.zeros((F, L * M), device="cuda", dtype=eng.dt)
    beta0.view(F, L, M).scatter_(2, idx0.unsqueeze(-1), sq)
    x = eng.ab(beta0)
    y = x + (P / args.snr) ** 0.5 * torch.randn((F, n), device="cuda", generator=gen, dtype=eng.dt)
    del beta0

    def step():
        return sparc_new._amp_device(eng, y, P, L, M, sq, t_max)[1]

    for _ in range(args.warmup):
        s = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        s = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    ser = float((s.view(F, L, M).argmax(2) != idx0).float().mean().item())
    flops = (2 * t_max - 1) * 2.0 * n * L * M * F            # A^T z every iteration, A beta from the second on
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["bf16_tflops_sustained"]) if os.path.exists(peaks) else 1400.0
    print(json.dumps({
        "workload": "dense-design AMP, L=%d M=%d n=%d (A %dx%d), t_max=%d, %d frames" % (L, M, n, n, L * M, t_max, F),
        "precision": args.precision, "ms_per_decode_batch": ms, "frames_per_s": F / (ms * 1e-3),
        "info_bits_per_s": F * L * logM / (ms * 1e-3), "section_error_rate": ser,
        "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                     "frac": flops / (ms * 1e-3) / 1e12 / peak,
                     "note": "algorithmic flops (one product per operand pair); the split-bf16 kernel issues 3x as many MMA flops"},
    }))


if __name__ == "__main__":
    main()

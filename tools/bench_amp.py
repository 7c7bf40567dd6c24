#!/usr/bin/env python
"""Throughput of the batched SPARC AMP decoder on BASELINE config 1 (sparc_public,
power-allocated, L=1024 M=32 R=1.5, sub-sampled DCT design, t_max=25, awgn_var=1).

Prints one JSON object: decoded info bits/s, frames/s, mean iterations, section error
rate, and the achieved fraction of the HBM state-streaming model of SURVEY 8(d)
(303 100 B per frame-iteration).  `python tools/bench_amp.py [--frames F] [--steps K]`.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BYTES_PER_FRAME_ITER = 303100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--atol", type=float, default=None, help="early-stop atol (default: the reference's 2e-15)")
    ap.add_argument("--config", default="c1", choices=["c1", "c4"],
                    help="c1: PA L=1024 M=32 R=1.5; c4: SC wave L=2048 M=512 R=1.5 omega=6 Lambda=32 t_max=40")
    args = ap.parse_args()
    import torch
    from ldpc_sparc_b200 import sparc

    global BYTES_PER_FRAME_ITER
    if args.config == "c1":
        cp = dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9)
        dp = dict(t_max=25)
    else:
        cp = dict(P=15.0, R=1.5, L=2048, M=512, spatially_coupled=True, omega=6, Lambda=32)
        dp = dict(t_max=40)
    if args.atol is not None:
        dp["atol"] = args.atol
    sparc.check_code_params(cp)
    sparc.check_decode_params(dp)
    L_, M_ = cp["L"], cp["M"]
    logM = int(np.log2(M_))
    W = sparc.create_base_matrix(**dict(cp, awgn_var=1.0))
    n = int(round(L_ * logM / cp["R"]))
    if W.ndim == 2:
        n = int(round(n / W.shape[0])) * W.shape[0]
    cp["n"] = n
    BYTES_PER_FRAME_ITER = (2 * L_ * M_ + 3 * n) * 4          # beta r+w, z r+w, y r in fp32
    Ab, Az = sparc.sparc_transforms(W, L_, M_, n, [1, 2])
    F = args.frames
    chunk = max(1, min(1024, (256 << 20) // (L_ * M_ * 4)))
    gen = torch.Generator(device="cuda").manual_seed(0)
    idx0 = torch.randint(0, M_, (F, L_), device="cuda", generator=gen, dtype=torch.int32)
    y = torch.empty((F, n), device="cuda", dtype=torch.float32)
    for f0 in range(0, F, chunk):     # x = A beta0 in chunks (beta0 is L*M*4 bytes per frame)
        f1 = min(F, f0 + chunk)
        beta0 = torch.zeros((f1 - f0, L_ * M_), device="cuda", dtype=torch.float32)
        beta0.view(f1 - f0, L_, M_).scatter_(2, idx0[f0:f1].long().unsqueeze(-1), 1.0)
        y[f0:f1] = Ab(beta0, precision="fp32") + torch.randn((f1 - f0, n), device="cuda", generator=gen)
    del beta0

    def step():
        return sparc.sparc_decode(y, cp, dp, 1.0, [1, 2], None, Ab, Az, precision=args.precision)

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    bits, idx, T, _, _ = out
    # end to end through the public call with host buffers: pinned y in, decided sections + t_final out
    e2e = None
    if args.config == "c1":
        h_y = torch.empty(y.shape, dtype=y.dtype).pin_memory()
        h_y.copy_(y)
        h_idx = torch.empty(idx.shape, dtype=idx.dtype).pin_memory()
        h_T = torch.empty(T.shape, dtype=T.dtype).pin_memory()

        def e2e_step():
            yd = h_y.to("cuda", non_blocking=True)
            o = sparc.sparc_decode(yd, cp, dp, 1.0, [1, 2], None, Ab, Az, precision=args.precision)
            h_idx.copy_(o[1], non_blocking=True)
            h_T.copy_(o[2], non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        import time
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        dt = (time.perf_counter() - t0) / args.steps
        assert torch.equal(h_idx.cuda(), idx)
        e2e = {"value": F * L_ * logM / dt, "unit": "info bits/s", "h2d_bytes_per_step": h_y.numel() * 4,
               "d2h_bytes_per_step": h_idx.numel() * 4 + h_T.numel() * 4, "ms_per_step": dt * 1e3}
    its = T.double().mean().item()
    frame_iters = float(T.sum().item())
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    achieved = BYTES_PER_FRAME_ITER * frame_iters / (ms * 1e-3) / 1e9
    print(json.dumps({
        "workload": "sparc_public AMP decode, config %s L=%d M=%d R=1.5 (n=%d), t_max=%d, %d frames sharing one design"
                    % (args.config, L_, M_, n, dp["t_max"], F),
        "precision": args.precision, "atol": dp.get("atol", 2e-15), "ms_per_step": ms, "frames_per_s": F / (ms * 1e-3),
        "info_bits_per_s": F * L_ * logM / (ms * 1e-3), "mean_iterations": its,
        "section_error_rate": float((idx != idx0).float().mean().item()), "e2e": e2e,
        "roofline": {"bound": "hbm (state-streaming model; really shared-memory/FFT bound, DESIGN.md)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak},
    }))


if __name__ == "__main__":
    main()

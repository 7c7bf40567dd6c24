#!/usr/bin/env python
"""Throughput of the complex-design AMP decoder (sub-sampled DFT, K-PSK sections) on a
config-1-sized code: L=1024, M=32, K=4, R=2.4 bits per complex channel use, power-allocated
(B=32), t_max=25.  Prints one JSON object (decode only, frames resident on the device)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--precision", default="fp32")
    args = ap.parse_args()
    import torch
    from ldpc_sparc_b200 import sparc
    cp = dict(P=15.0, R=2.4, L=1024, M=32, complex=True, modulated=True, K=4, power_allocated=True, B=32, R_PA_ratio=0.9)
    dp = dict(t_max=25)
    bits, beta0, x, Ab, Az = sparc.sparc_encode(cp, 1.0, [1, 2])
    sparc.check_decode_params(dp)
    L, M, K, n = cp["L"], cp["M"], cp["K"], cp["n"]
    F = args.frames
    gen = torch.Generator(device="cuda").manual_seed(0)
    sym0 = torch.randint(0, K * M, (F, L), device="cuda", generator=gen, dtype=torch.int32)
    dt = torch.complex64 if args.precision == "fp32" else torch.complex128
    con = torch.as_tensor(np.asarray(sparc.psk_constel(K)), device="cuda").to(dt)
    b0 = torch.zeros((F, L * M), dtype=dt, device="cuda")
    b0.view(F, L, M).scatter_(2, (sym0 % M).long().unsqueeze(-1), con[(sym0 // M).long()].unsqueeze(-1))
    xs = Ab(b0, args.precision)
    del b0
    noise = torch.view_as_complex(torch.randn((F, n, 2), device="cuda", generator=gen, dtype=xs.real.dtype))
    y = xs + float(np.sqrt(0.5)) * noise

    def step():
        return sparc.sparc_decode(y, cp, dp, 1.0, [1, 2], sym0, Ab, Az, precision=args.precision)

    out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"workload": "complex-design AMP, L=1024 M=32 K=4 R=2.4 PA (n=%d complex, w=%d), t_max=25, %d frames"
                                  % (n, Ab.design.w, F), "precision": args.precision, "ms_per_step": ms,
                      "frames_per_s": F / (ms * 1e-3), "info_bits_per_s": F * L * 7 / (ms * 1e-3),
                      "mean_iterations": float(out[2].float().mean().item()),
                      "section_error_rate": float((out[1] != sym0).float().mean().item())}))


if __name__ == "__main__":
    main()

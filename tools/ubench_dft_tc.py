#!/usr/bin/env python
"""Measured cost of a tensor-core formulation of the AMP transforms.

A length-4096 transform factors as 64 x 64: two stages of 64 DFT-64s with a twiddle and a transposition between
them.  One stage over a batch is the GEMM  out[128, 64 T] = D64 [128 x 128] x X [128, 64 T]  (D64 = the DFT-64
matrix in real form, X = real and imaginary parts stacked), T = transforms in the batch.  This script runs exactly
that GEMM through the library's tcgen05 kernel (csrc/tc_gemm.cuh, split-fp16 operands: three MMAs per product, the
precision the 1e-4 parity bound needs) and reports DFT-64 STAGES of 4096-point transforms per second, next to the
rate the shared-memory FFT kernel sustains inside the AMP decoder:

    amp_pair_kernel, config 1: 57.3 k frames/s x 23.84 iterations x 64 transforms = 87.4 M transforms/s,

each of which contains the work of TWO such stages plus scatter / gather, the real-FFT split and the denoiser.
Usage: python tools/ubench_dft_tc.py [--transforms T] [--steps K]      (prints one JSON object)"""
import argparse
import ctypes as ct
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transforms", type=int, default=16384, help="4096-point transforms per launch (64 columns each)")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from ldpc_sparc_b200 import _lib
    _lib.require_device()
    L_ = _lib.lib()
    k = np.arange(64)
    ang = -2 * np.pi * np.outer(k, k) / 64
    C, S = np.cos(ang), np.sin(ang)
    D = np.block([[C, -S], [S, C]])                      # [re; im] -> [re; im] of the DFT-64
    h = ct.c_void_p()
    _lib.check(L_.lsb_tc_design_create(np.ascontiguousarray(D).ctypes.data_as(_lib.c_f64p), 128, 128, ct.byref(h)), "create")
    cols = 64 * args.transforms
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((cols, 128), device="cuda", generator=gen)
    out = torch.empty((cols, 128), device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def step():
        _lib.check(L_.lsb_tc_ab_dev_f32(h, x.data_ptr(), out.data_ptr(), cols, st), "ab")

    step()
    torch.cuda.synchronize()
    xs = x[:256].double().cpu().numpy()
    ref = np.fft.fft(xs[:, :64] + 1j * xs[:, 64:], axis=1)
    got = out[:256].double().cpu().numpy()
    err = np.max(np.abs(got[:, :64] + 1j * got[:, 64:] - ref)) / np.max(np.abs(ref))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    stages = args.transforms / (ms * 1e-3)
    fft_transforms = 57.3e3 * 23.84 * 64
    print(json.dumps({
        "what": "one DFT-64 stage of T 4096-point transforms as a tcgen05 GEMM [128 x 128] x [128, 64 T], split-fp16 (3 MMAs)",
        "transforms_per_launch": args.transforms, "ms_per_launch": ms, "max_rel_err_vs_numpy_fft": err,
        "dft64_stages_per_s": stages,
        "upper_bound_transforms_per_s": stages / 2,
        "note": "includes the fp32 -> fp16 hi/lo split of the operand (split8_kernel) and the fp32 store of the result; "
                "a 4096-point transform needs two stages + twiddles + a transposition, so stages/2 is an upper bound",
        "bytes_per_stage_transform": 64 * (128 * 4 + 2 * 128 * 2 * 2 + 128 * 4),
        "fft_kernel_transforms_per_s_inside_amp": fft_transforms,
        "ratio_upper_bound_over_fft": stages / 2 / fft_transforms}))
    L_.lsb_tc_design_destroy(h)


if __name__ == "__main__":
    main()

"""Secondary workloads of BASELINE.json (configs 1, 3, 4, 5), timed with the headline's rules on every rank.

Each function runs on ALL ranks of a torchrun job (one process per GPU, weak scaling: the per-GPU batch is
fixed), does `warmup` untimed steps, times exactly `steps` steps between barriers with CUDA events on the launch
stream, takes the max over ranks, samples the SM clocks during the timed region, and returns one JSON-able dict
whose `value` is the whole-job aggregate.  bench.py calls them after the headline; the tools/bench_*.py command
lines call them with a single-rank context.
"""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop, self.thread = index, [], threading.Event(), None

    def _loop(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([t.strip() for t in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Ctx:
    """Rank context: one process per GPU.  `dist` is torch.distributed when world > 1."""

    def __init__(self, rank=0, local_rank=0, world=1):
        import torch
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.torch = torch
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, xs):
        t = self.torch.tensor([float(x) for x in xs], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def timed(ctx, step, steps, warmup, host_sync=False):
    """warmup untimed steps, then `steps` steps between barriers.  Device time from CUDA events on the current
    stream (or, with host_sync, wall time of calls that return when their outputs are on the host), max over
    ranks.  Returns (total_ms, clocks summary)."""
    torch = ctx.torch
    for _ in range(warmup):
        step()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(ctx.local_rank) as clocks:
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        ctx.barrier()
    ms = wall_ms if host_sync else e0.elapsed_time(e1)
    return ctx.max_over_ranks(ms), clocks.summary()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ SPARC AMP, DCT design (configs 1 and 4)
def amp_code(config):
    from ldpc_sparc_b200 import sparc
    if config == "c1":
        cp = dict(P=15.0, R=1.5, L=1024, M=32, power_allocated=True, B=32, R_PA_ratio=0.9)
        dp = dict(t_max=25)
    else:
        cp = dict(P=15.0, R=1.5, L=2048, M=512, spatially_coupled=True, omega=6, Lambda=32)
        dp = dict(t_max=40)
    sparc.check_code_params(cp)
    sparc.check_decode_params(dp)
    L_, M_ = cp["L"], cp["M"]
    logM = int(np.log2(M_))
    W = sparc.create_base_matrix(**dict(cp, awgn_var=1.0))
    n = int(round(L_ * logM / cp["R"]))
    if W.ndim == 2:
        n = int(round(n / W.shape[0])) * W.shape[0]
    cp["n"] = n
    Ab, Az = sparc.sparc_transforms(W, L_, M_, n, [1, 2])
    return cp, dp, Ab, Az, logM


def amp_workload(ctx, config="c1", frames=16384, steps=5, warmup=2, precision="fp32", atol=None, e2e=True):
    """sparc_public AMP decode, `frames` frames per GPU sharing one design (messages and noise per frame,
    generated on the device).  value: inputs resident in HBM.  e2e: sparc.sparc_decode_host =
    lsb_sparc_amp_host_f32 with pinned host buffers (y in; decisions, t_final, psi out)."""
    torch = ctx.torch
    from ldpc_sparc_b200 import sparc
    cp, dp, Ab, Az, logM = amp_code(config)
    if atol is not None:
        dp["atol"] = atol
    L_, M_, n = cp["L"], cp["M"], cp["n"]
    bytes_per_frame_iter = (2 * L_ * M_ + 3 * n) * 4          # beta r+w, z r+w, y r in fp32
    F = frames
    chunk = max(1, min(1024, (256 << 20) // (L_ * M_ * 4)))
    gen = torch.Generator(device=ctx.dev).manual_seed(1000 + ctx.rank)
    idx0 = torch.randint(0, M_, (F, L_), device=ctx.dev, generator=gen, dtype=torch.int32)
    y = torch.empty((F, n), device=ctx.dev, dtype=torch.float32)
    for f0 in range(0, F, chunk):     # x = A beta0 in chunks (beta0 is L*M*4 bytes per frame)
        f1 = min(F, f0 + chunk)
        beta0 = torch.zeros((f1 - f0, L_ * M_), device=ctx.dev, dtype=torch.float32)
        beta0.view(f1 - f0, L_, M_).scatter_(2, idx0[f0:f1].long().unsqueeze(-1), 1.0)
        y[f0:f1] = Ab(beta0, precision="fp32") + torch.randn((f1 - f0, n), device=ctx.dev, generator=gen)
    del beta0
    res = {}

    def step():
        res["out"] = sparc.sparc_decode(y, cp, dp, 1.0, [1, 2], None, Ab, Az, precision=precision)

    total_ms, clocks = timed(ctx, step, steps, warmup)
    ms = total_ms / steps
    bits, idx, T, _, _ = res["out"]
    frame_iters_rank = float(T.sum().item())
    sec_err, frame_iters, n_frames = ctx.sum_over_ranks([float((idx != idx0).sum().item()), frame_iters_rank, F])
    hbm, tens, src = peaks()
    achieved = bytes_per_frame_iter * frame_iters_rank / (ms * 1e-3) / 1e9      # this rank's kernel
    out = {
        "workload": "sparc_public AMP decode, config %s L=%d M=%d R=1.5 (n=%d), t_max=%d, %d frames per GPU sharing one design"
                    % (config, L_, M_, n, dp["t_max"], F),
        "metric": "decoded info bits/s", "value": ctx.world * F * L_ * logM / (ms * 1e-3), "unit": "info bits/s",
        "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "scaling": "weak", "dtype": "f32",
        "precision": precision, "atol": dp.get("atol", 2e-15),
        "frames_per_s": ctx.world * F / (ms * 1e-3), "mean_iterations": frame_iters / n_frames,
        "section_error_rate": sec_err / (n_frames * L_),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "peak_source": src, "traffic": traffic_entry("amp_%s" % config),
                     "model": "%d B per frame-iteration (state-streaming model of SURVEY 8(d): beta r+w, z r+w, y r); "
                              "the FFT kernels are latency-bound under their shared-memory pipe (roofline_smem), not HBM-bound (DESIGN.md 4)"
                              % bytes_per_frame_iter},
        "clocks": clocks,
    }
    if precision == "fp32" and config == "c1":
        # the pipe the FFT kernel keeps busiest: shared-memory wavefronts.  100.2 k wavefronts per frame-iteration
        # (ncu l1tex__data_pipe_lsu_wavefronts_mem_shared.sum of amp_pair_kernel<11> over 8192 frames x 23.86
        # iterations, profiles/r2_amp_pair_v7_ncu_full.csv; 125.7 k before the fused passes) against one wavefront
        # per SM per clock
        wf = 100.2e3 * frame_iters_rank / (ms * 1e-3)
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        peak_wf = float(torch.cuda.get_device_properties(ctx.dev).multi_processor_count) * float(sm_mhz) * 1e6
        out["roofline_smem"] = {"bound": "shared memory", "achieved": wf / 1e9, "peak": peak_wf / 1e9,
                                "unit": "G wavefronts/s", "frac": wf / peak_wf,
                                "wavefronts_per_frame_iteration": 100.2e3,
                                "note": "64 length-4096 transforms per frame-iteration through a 32 KB shared-memory buffer; "
                                        "15 % of the wavefronts are bank conflicts of the random gathers; the kernel is "
                                        "latency-bound below this pipe's peak (2 CTAs per SM in barrier lockstep; DESIGN.md 4)"}
    if precision == "tc":
        # explicit block-banded design on the tensor cores (csrc/sparc_band_tc.cu): the bound is the tcgen05 pipe.
        # Algorithmic flops: 2 per stored design entry per product; a frame that runs T iterations does T products
        # A^T z and T - 1 products A beta.  The split-fp16 kernel issues 3x that (plus tile padding).
        d = Ab.design
        nnz = float(np.count_nonzero(d.Wm)) * d.Mr * d.Mc
        flops = 2.0 * nnz * (2.0 * frame_iters_rank - F)
        ach = flops / (ms * 1e-3) / 1e12
        out["dtype"] = "f16x3 (split) products, f32 elsewhere"
        out["roofline_hbm_model"] = out["roofline"]
        out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tens, "unit": "TFLOP/s", "frac": ach / tens,
                           "peak_source": src, "traffic": None,
                           "model": "%.2f GFLOP per frame-iteration (2 flops per stored design entry and product, %d non-zero "
                                    "blocks of %d x %d; A beta skipped in a frame's first iteration); algorithmic flops -- the "
                                    "split-fp16 kernel issues 3x" % (4.0 * nnz / 1e9, int(np.count_nonzero(d.Wm)), d.Mr, d.Mc)}
        out["design_bytes"] = int(_lib_bytes(d))
    if e2e and precision in ("fp32", "tc"):
        h_y = torch.empty(y.shape, dtype=torch.float32).pin_memory()
        h_y.copy_(y)
        h_idx = torch.empty(idx.shape, dtype=torch.int32).pin_memory()
        h_T = torch.empty(T.shape, dtype=torch.int32).pin_memory()
        h_psi = torch.empty((F, Ab.design.Lc), dtype=torch.float64).pin_memory()

        def e2e_step():
            sparc.sparc_decode_host(h_y, cp, dp, 1.0, Ab, out=(h_idx, h_T, h_psi), precision=precision)

        n_e2e = max(1, min(steps, 3))
        e_ms, _ = timed(ctx, e2e_step, n_e2e, 1, host_sync=True)
        assert torch.equal(h_idx.to(ctx.dev), idx), "host-entry decisions differ from the device path"
        out["e2e"] = {"value": ctx.world * F * L_ * logM / (e_ms / n_e2e * 1e-3), "unit": "info bits/s",
                      "h2d_bytes_per_step": h_y.numel() * 4,
                      "d2h_bytes_per_step": h_idx.numel() * 4 + h_T.numel() * 4 + h_psi.numel() * 8,
                      "ms_per_step": e_ms / n_e2e,
                      "api": "%s (C-ABI, pinned host y in; decisions, t_final, psi out)"
                             % ("lsb_band_amp_host_f32" if precision == "tc" else "lsb_sparc_amp_host_f32")}
    else:
        out["e2e"] = None
    return out


def _lib_bytes(design):
    from ldpc_sparc_b200 import _lib
    return _lib.lib().lsb_band_design_bytes(design.band_handle())


def traffic_entry(key):
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(key + "_dram_bytes_per_launch")
    return None


# ------------------------------------------------------------------ dense design AMP (config 3)
def dense_workload(ctx, frames=1024, steps=5, warmup=2, precision="tc", L=512, M=64, snr=6.0):
    """sparc_sophie dense Gaussian design, L=512 M=64, R=1 -> n=3072, A 3072 x 32768, P=19.44, t_max=25;
    `frames` frames per GPU share the design.  Tensor roofline of SURVEY 8(d): 402.7 MFLOP per frame-iteration."""
    torch = ctx.torch
    from ldpc_sparc_b200 import sparc_new
    P, t_max = 19.44, 25
    logM = int(np.log2(M))
    n = L * logM
    A = sparc_new.create_design_matrix(L, M, n, 0)
    eng = sparc_new._engine(A, precision)
    F = frames
    gen = torch.Generator(device=ctx.dev).manual_seed(2000 + ctx.rank)
    idx0 = torch.randint(0, M, (F, L), device=ctx.dev, generator=gen)
    sq = float(np.sqrt(n * P / L))
    beta0 = torch.zeros((F, L * M), device=ctx.dev, dtype=eng.dt)
    beta0.view(F, L, M).scatter_(2, idx0.unsqueeze(-1), sq)
    x = eng.ab(beta0)
    y = x + (P / snr) ** 0.5 * torch.randn((F, n), device=ctx.dev, generator=gen, dtype=eng.dt)
    del beta0
    res = {}

    def step():
        res["s"] = sparc_new._amp_device(eng, y, P, L, M, sq, t_max)[1]

    total_ms, clocks = timed(ctx, step, steps, warmup)
    ms = total_ms / steps
    ser_sum, = ctx.sum_over_ranks([float((res["s"].view(F, L, M).argmax(2) != idx0).sum().item())])
    flops = (2 * t_max - 1) * 2.0 * n * L * M * F            # A^T z every iteration, A beta from the second on (per rank)
    _, tf, src = peaks()
    ach = flops / (ms * 1e-3) / 1e12
    e2e = None
    if precision == "tc":
        # the same decode through the host-pointer C-ABI entry: pinned y in, MAP section decisions out
        h_y = torch.empty(y.shape, dtype=torch.float32).pin_memory()
        h_y.copy_(y)
        h_idx = torch.empty((F, L), dtype=torch.int32).pin_memory()
        sp, dp = dict(P=P, L=L, M=M), dict(t_max=t_max)

        def e2e_step():
            sparc_new.sparc_amp_decide_host(h_y, sp, dp, A, out=h_idx)

        n_e2e = max(1, min(steps, 3))
        e_ms, _ = timed(ctx, e2e_step, n_e2e, 1, host_sync=True)
        assert torch.equal(h_idx.to(ctx.dev).long(), res["s"].view(F, L, M).argmax(2)), "host-entry decisions differ"
        e2e = {"value": ctx.world * F * L * logM / (e_ms / n_e2e * 1e-3), "unit": "info bits/s",
               "h2d_bytes_per_step": h_y.numel() * 4, "d2h_bytes_per_step": h_idx.numel() * 4, "ms_per_step": e_ms / n_e2e,
               "api": "lsb_tc_amp_host_f32 (C-ABI, pinned host y in; MAP section decisions out)"}
    return {
        "workload": "dense-design AMP, L=%d M=%d n=%d (A %dx%d), t_max=%d, %d frames per GPU" % (L, M, n, n, L * M, t_max, F),
        "metric": "decoded info bits/s", "value": ctx.world * F * L * logM / (ms * 1e-3), "unit": "info bits/s",
        "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "scaling": "weak",
        "dtype": "f16x3 (split) products, f32 elsewhere" if precision == "tc" else precision, "precision": precision,
        "frames_per_s": ctx.world * F / (ms * 1e-3), "section_error_rate": ser_sum / (ctx.world * F * L),
        "roofline": {"bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf, "peak_source": src,
                     "traffic": None,
                     "model": "402.7 MFLOP per frame-iteration at L=512 M=64 (SURVEY 8(d)); algorithmic flops, one product per "
                              "operand pair -- the split-fp16 kernel issues 3x as many MMA flops"},
        "e2e": e2e, "clocks": clocks,
    }


# ------------------------------------------------------------------ LDPC-protected dense-design SPARC (config 5)
def c5_workload(ctx, frames=256, steps=5, warmup=2, precision="tc", small=False, sims=("SPARC+LDPC", "Integrated_decoder"),
                snrs=(2.0, 6.0)):
    """BASELINE config 5 (performance_plots_general.py:21-47): 802.16 r1/2 z=150, mults=4, M=2 -> L = n = 14400,
    A 14400 x 28800, t_max = 25.  One step = `frames` frames per GPU per SNR point and scheme, generated, decoded and
    counted on the GPU.  Roofline: GEMM flops 2 (2 n LM) per frame-iteration against the sustained bf16
    peak, and the BP state-streaming bytes (2 Nmsg + 2 Nv) 4 B per block-iteration for the record."""
    torch = ctx.torch
    from ldpc_sparc_b200 import ldpc, sparc_new, sparc_sim_new
    from ldpc_sparc_b200.param_calc import param_calc
    mults = 1 if small else 4
    overall, L_sparc, L_sl, lengths = param_calc(mults, 1, '802.16', '1/2', 1 / 2, 150, 1)
    sp = {'P': 19.44, 'R': overall, 'L': L_sparc, 'M': 2}
    spl = {'P': 19.44, 'R': 1, 'L': L_sl, 'M': 2}
    lp = {'standard': '802.16', 'rate': '1/2', 'z': 150, 'int_rate': 1 / 2, 'mults': mults}
    t_max = 25
    res = {}
    it = {"n": 0}

    def step():
        it["n"] += 1
        res["out"] = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': t_max}, list(snrs), frames * ctx.world,
                                                     sims=tuple(sims), seed=it["n"], chunk=frames, precision=precision)
        torch.cuda.synchronize()

    total_ms, clocks = timed(ctx, step, steps, warmup, host_sync=True)
    ms = total_ms / steps
    out = res["out"]
    n_frames = frames * len(snrs) * len(sims) * ctx.world              # frames decoded per step, whole job
    c = ldpc.code('802.16', '1/2', 150)
    n, LM = L_sl, L_sl * 2
    gemm_flops = (2 * t_max - 1) * 2.0 * n * LM * (n_frames / ctx.world)       # per rank and step
    _, tf, src = peaks()
    hbm, _, _ = peaks()
    ach = gemm_flops / (ms * 1e-3) / 1e12
    return {
        "workload": "config 5: 802.16 r1/2 z=150 x %d blocks, M=2, L=n=%d, t_max=%d, %d frames per GPU per point, SNR %s, schemes %s"
                    % (mults, L_sl, t_max, frames, list(snrs), list(sims)),
        "metric": "decoded user bits/s", "value": float(out["bits"].sum()) / (ms * 1e-3), "unit": "info bits/s",
        "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "scaling": "weak",
        "dtype": "f16x3 (split) products, f32 elsewhere", "precision": precision,
        "frames_per_s": n_frames / (ms * 1e-3),
        "ber": {name: [float(v) for v in out["ber_store_averages"][i]] for i, name in enumerate(sims)},
        "roofline": {"bound": "tensor", "achieved": ach, "peak": tf, "unit": "TFLOP/s", "frac": ach / tf, "peak_source": src,
                     "traffic": None,
                     "model": "2 (2 n LM) = %.2f GFLOP per frame-iteration at L=n=%d (SURVEY 8(d)), (2 t_max - 1)/2 such per "
                              "frame; algorithmic flops (the split-fp16 kernel issues 3x). BP: (2 Nmsg + 2 Nv) 4 B = %d B per "
                              "block-iteration, %d blocks per frame, hidden under the GEMMs' time"
                              % (4.0 * n * LM / 1e9, n, (2 * int(c.Nmsg) + 2 * int(c.N)) * 4, mults),
                     "includes": "frame generation (encode, A beta0, noise) and error counting inside the step"},
        "e2e": None, "clocks": clocks,
    }


def single_rank_ctx():
    import torch
    torch.cuda.set_device(0)
    return Ctx(0, 0, 1)

#!/usr/bin/env python
"""Headline benchmark: decoded info bits/s of the batched LDPC BP decoder.

    python bench.py --gpus N --steps K --warmup W          (this repo's CUDA path)
    python bench.py --impl reference --gpus N ...          (reference CPU path, host cores)

Workload (BASELINE.json configs[1]): ldpc_jossy 802.11n rate-1/2 z=81 (n=1944)
sum-product BP (sumprod2), max 50 iterations, 65 536 AWGN frames per GPU per step
(random codewords, BPSK, Es/N0 = 1.2 dB: the lowest of SURVEY section 8(d)'s three
operating points, i.e. the most iterations per frame).  One step = one decode of
the whole batch.  Frames are sharded over ranks with no data-path collective
(weak scaling: 65 536 frames per GPU); the only collective is the final
all-reduce of the error / iteration counters and the max over ranks of the time.

Prints ONE JSON line (see the task contract): value = whole-job info bits/s with
inputs resident in HBM; e2e = same metric through the C-ABI host-pointer call with
pinned host buffers (H2D + D2H inside the timed region); roofline = achieved
algorithmic GB/s of the dominant kernel (71 280 B per frame-iteration, SURVEY
8(d)) over the measured HBM peak; cpu_baseline = the reference's own c_ldpc.c
(oracle/_ref) on all host cores over a bounded sample of the same frames.
"""
import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

N_FRAMES = 65536
MAX_IT = 50
SNR_DB = 1.2
# warp instructions bp_qc_kernel<float, sumprod2, 8> executes per frame-iteration on this workload: ncu
# smsp__inst_executed.sum / frame-iterations of the profiled launch (profiles/r2_bp_edomain_v2_ncu_full.csv)
WARP_INSTR_PER_FRAME_ITER = 10632.0      # 18 166 078 038 / 1 708 694
BYTES_PER_FRAME_ITER = 71280          # (2*Nmsg + 2*Nv) * 4 B, SURVEY.md section 8(d)
METRIC = "decoded info bits/s (LDPC 802.11n r1/2 n=1944 sum-product BP, <=50 its, batched AWGN frames)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=N_FRAMES, help="frames per GPU per step")
    ap.add_argument("--snr", type=float, default=SNR_DB)
    ap.add_argument("--max-it", type=int, default=MAX_IT)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target wall time of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mc", action="store_true", help="skip the fused generate-decode-count Monte-Carlo leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the fp64 re-decode of the batch after the timed region")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary workloads (SPARC AMP config 1, dense-design AMP config 3)")
    return ap.parse_args()


def config_dict(args, world):
    return {
        "workload": "ldpc_jossy 802.11n r1/2 z=81 (n=1944, K=972) sumprod2 BP, max %d its, %d AWGN frames per GPU "
                    "per step at Es/N0=%.1f dB (BASELINE configs[1])" % (args.max_it, args.frames, args.snr),
        "frames_per_gpu": args.frames,
        "max_it": args.max_it,
        "esn0_db": args.snr,
        "sharding": "frames/%d ranks, no data-path collective" % world,
        "l2_policy": "inputs larger than L2 (%.0f MB LLR batch per GPU vs 126 MB L2)" % (args.frames * 1944 * 4 / 1e6),
    }


# ------------------------------------------------------------------ CPU baseline
def host_frames(code, n, snr_db, seed):
    rng = np.random.RandomState(seed)
    x = code.encode(rng.randint(0, 2, (n, code.K)))
    s2 = 10 ** (-snr_db / 10)
    return x, 2 / s2 * ((1 - 2.0 * x) + np.sqrt(s2) * rng.randn(n, code.N))


def cpu_decode_threads(code, llr, max_it, threads):
    """Decode rows of llr with the compiled reference (oracle/_ref/c_ldpc_ref.so) or,
    if absent, the oracle port, on `threads` host threads (ctypes releases the GIL).
    Returns (seconds, kind, iters)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ldpc_oracle as orc
    g = orc.Graph(code.vdeg, code.cdeg, code.intrlv)
    kind = "reference" if orc.ref_lib() is not None else "port"
    impl = "ref" if kind == "reference" else "oracle"
    orc.decode(g, llr[0], 2, impl=impl)        # load the library outside the timed region
    iters = np.zeros(len(llr), dtype=np.int32)
    chunks = np.array_split(np.arange(len(llr)), threads)

    def work(idx):
        for f in idx:
            _, iters[f] = orc.decode(g, llr[f], max_it, impl=impl)

    ts = [threading.Thread(target=work, args=(c,)) for c in chunks if len(c)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0, kind, iters


def cpu_baseline(code, args, seconds):
    cores = os.cpu_count() or 1
    probe_n = max(cores, 8)
    _, llr = host_frames(code, probe_n, args.snr, 12345)
    dt, kind, _ = cpu_decode_threads(code, llr, args.max_it, cores)
    per_round = dt                               # probe_n frames on all cores
    n = int(max(probe_n, min(args.frames, probe_n * seconds / max(per_round, 1e-3))))
    _, llr = host_frames(code, n, args.snr, 777)
    dt, kind, iters = cpu_decode_threads(code, llr, args.max_it, cores)
    return {
        "value": n * code.K / dt,
        "unit": "info bits/s",
        "cores": cores,
        "kind": kind,
        "sample": "%d frames of the same workload (Es/N0 %.1f dB, max %d its, mean its %.1f), decode only, "
                  "%d threads x c_ldpc.c -O2, %.1f s" % (n, args.snr, args.max_it, float(iters.mean()), cores, dt),
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on the host cores."""
    if rank != 0:
        return
    from ldpc_sparc_b200 import ldpc
    code = ldpc.code("802.11n", "1/2", 81)
    cores = os.cpu_count() or 1
    total_steps = args.steps + args.warmup
    per_step_s = max(2.0, min(20.0, 150.0 / max(total_steps, 1)))
    _, llr = host_frames(code, max(cores, 8), args.snr, 12345)
    dt, kind, _ = cpu_decode_threads(code, llr, args.max_it, cores)
    n = int(max(cores, len(llr) * per_step_s / max(dt, 1e-3)))
    _, llr = host_frames(code, n, args.snr, 777)
    for _ in range(args.warmup):
        cpu_decode_threads(code, llr, args.max_it, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, kind, iters = cpu_decode_threads(code, llr, args.max_it, cores)
        t += dt
    value = n * code.K * args.steps / t
    sample = "%d frames per step (bounded sample of the %d-frame workload), %d host threads" % (n, args.frames, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "info bits/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world),
        "cpu_baseline": {"value": value, "unit": "info bits/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "info bits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------- main
def bind_to_gpu_numa(torch, local_rank):
    """Multi-rank runs: keep this process (and so its pinned host buffers, first-touched by it) on the
    CPUs NVML reports as local to its GPU; with 8 ranks the host<->device copies of the e2e leg
    otherwise cross the socket interconnect.  Best effort -- any failure leaves the affinity alone."""
    try:
        import pynvml
        p = torch.cuda.get_device_properties(local_rank)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)).encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import workloads
    from workloads import ClockSampler
    from ldpc_sparc_b200 import _lib, ldpc

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        bind_to_gpu_numa(torch, local_rank)
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    code = ldpc.code("802.11n", "1/2", 81)
    F, N, K = args.frames, code.N, code.K

    # synthetic frames: 1024 random codewords (host encoder) tiled over the batch,
    # independent noise per frame from a per-rank Philox stream
    x, _ = host_frames(code, 1024, args.snr, 99)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    s2 = 10 ** (-args.snr / 10)
    sgn = torch.from_numpy(1.0 - 2.0 * x).float().to(dev)
    llr = torch.empty((F, N), dtype=torch.float32, device=dev)
    for f0 in range(0, F, 8192):
        f1 = min(F, f0 + 8192)
        idx = torch.arange(f0, f1, device=dev) % 1024
        llr[f0:f1] = (2 / s2) * (sgn[idx] + (s2 ** 0.5) * torch.randn((f1 - f0, N), device=dev, generator=gen))
    truth = (sgn[torch.arange(F, device=dev) % 1024] < 0)
    app = torch.empty_like(llr)
    iters = torch.empty(F, dtype=torch.int32, device=dev)
    g = code._handle()
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        _lib.check(L.lsb_ldpc_decode_dev_f32(g, llr.data_ptr(), app.data_ptr(), None, iters.data_ptr(), F,
                                             args.max_it, 0, 1.0, stream), "decode")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    L.lsb_launch_count_reset()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        ev[0].record()
        for i in range(args.steps):
            step()
            ev[i + 1].record()
        barrier()
    launches = int(L.lsb_launch_count())
    total_ms = ev[0].elapsed_time(ev[-1])
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * F * K * args.steps / (total_ms * 1e-3)

    # counters (the Monte-Carlo result of the step): one all-reduce, not in the data path
    executed = torch.clamp(iters.long() + 1, max=args.max_it)
    hard_err = ((app < 0) != truth)
    counters = torch.stack([hard_err.sum(), hard_err.any(1).sum(), torch.tensor(F, device=dev),
                            iters.long().sum(), executed.sum()]).to(torch.int64)
    if world > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    bit_errs, frame_errs, frames, it_sum, exec_sum = (int(v) for v in counters.tolist())

    # roofline of the dominant kernel (bp_qc_kernel<float, SUMPROD2, 8>): this rank's launch
    kernel_ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    kernel_ms_avg = ev[0].elapsed_time(ev[-1]) / args.steps
    frame_iters_rank = int(executed.sum().item())
    achieved = BYTES_PER_FRAME_ITER * frame_iters_rank / (kernel_ms_avg * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("bp_qc_kernel_f32_dram_bytes_per_launch")

    out = {
        "metric": METRIC, "value": value, "unit": "info bits/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, world),
        "gpu_launches": launches,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "kernel": "bp_qc_kernel<float,SUMPROD2,8>",
            "kernel_ms_avg": kernel_ms_avg, "kernel_ms_min": kernel_ms,
            "model": "71280 B per frame-iteration (state-streaming model, SURVEY 8(d)) x %d frame-iterations "
                     "executed per launch; messages are shared-memory resident, so real DRAM traffic is only "
                     "LLR-in + APP-out (15 552 B/frame): the true limiter is the MUFU pipe (DESIGN.md)" % frame_iters_rank,
            "frames_per_s_per_gpu": F / (kernel_ms_avg * 1e-3),
        },
        "monte_carlo": {"frames": frames, "frame_errors": frame_errs, "bit_errors": bit_errs,
                        "mean_iterations": it_sum / max(frames, 1), "mean_executed_iterations": exec_sum / max(frames, 1)},
    }

    # the same step with the frames GENERATED inside the kernel (lsb_ldpc_mc_dev_f32: Philox noise + BPSK + LLR scaling in
    # the load phase, error counting in the epilogue): what a Monte-Carlo sweep runs, no 510 MB LLR batch in HBM
    if not args.no_mc:
        cw_pool = (sgn < 0).to(torch.uint8).contiguous()
        mc_counters = torch.zeros(4, dtype=torch.int64, device=dev)

        def mc_step(i=[0]):
            i[0] += 1
            code.mc_decode(cw_pool, F, s2, seed=4242, frame0=(i[0] * world + rank) * F, max_itcount=args.max_it,
                           counters=mc_counters)

        mc_ms, _ = workloads.timed(workloads.Ctx(rank, local_rank, world), mc_step, args.steps, 2)
        mc = mc_counters.clone()
        if world > 1:
            dist.all_reduce(mc, op=dist.ReduceOp.SUM)
        mcl = [int(v) for v in mc.tolist()]
        out["monte_carlo_fused"] = {
            "value": world * F * K * args.steps / (mc_ms * 1e-3), "unit": "info bits/s", "ms_per_step": mc_ms / args.steps,
            "frames_per_s": world * F * args.steps / (mc_ms * 1e-3), "frames": mcl[2], "frame_errors": mcl[1],
            "bit_errors": mcl[0], "mean_iterations": mcl[3] / max(mcl[2], 1),
            "api": "lsb_ldpc_mc_dev_f32 (generate -> decode -> count in one kernel; warm-up steps included in the counters)"}

    # parity at bench scale (outside the timed region): the same frames through the fp64 kernel, whose iteration
    # counts and decisions are the reference's (tests/test_ldpc_gpu.py: exact against c_ldpc.c)
    if not args.no_parity:
        llr64 = llr.double()
        app64 = torch.empty_like(llr64)
        it64 = torch.empty(F, dtype=torch.int32, device=dev)
        _lib.check(L.lsb_ldpc_decode_dev_f64(g, llr64.data_ptr(), app64.data_ptr(), None, it64.data_ptr(), F,
                                             args.max_it, 0, 1.0, stream), "decode f64")
        torch.cuda.synchronize()
        hard64, hard32 = app64 < 0, app < 0
        err64, err32 = (hard64 != truth).any(1), hard_err.any(1)
        par = torch.stack([err32.sum(), err64.sum(), (err32 != err64).sum(), (hard64 != hard32).any(1).sum(),
                           (it64 != iters).sum(), ((hard64 != hard32).any(1) & (it64 < args.max_it)).sum(),
                           (hard64 != truth).sum()]).to(torch.int64)
        if world > 1:
            dist.all_reduce(par, op=dist.ReduceOp.SUM)
        fe32, fe64, d_status, d_dec, d_it, d_dec_conv, be64 = (int(v) for v in par.tolist())
        out["parity_at_bench_scale"] = {
            "frames": frames, "frame_errors_fp32": fe32, "frame_errors_fp64": fe64, "bit_errors_fp32": bit_errs,
            "bit_errors_fp64": be64, "frames_with_different_error_status": d_status,
            "frames_with_different_hard_decisions": d_dec, "of_which_fp64_converged": d_dec_conv,
            "frames_with_different_iteration_counts": d_it,
            "note": "fp64 kernel = the reference's arithmetic (bit-exact iteration counts and decisions against c_ldpc.c in "
                    "the tests); fp32 differences are confined to frames the reference does not converge on within max_it",
        }
        del llr64, app64, hard64
        torch.cuda.empty_cache()

    # end to end through the C-ABI host-pointer call (pinned host buffers)
    if not args.no_e2e:
        h_llr = torch.empty((F, N), dtype=torch.float32).pin_memory()
        h_llr.copy_(llr)
        h_app = torch.empty((F, N), dtype=torch.float32).pin_memory()
        h_it = torch.empty(F, dtype=torch.int32).pin_memory()

        def e2e_step():
            _lib.check(L.lsb_ldpc_decode_host_f32(g, h_llr.data_ptr(), h_app.data_ptr(), None, h_it.data_ptr(), F,
                                                  args.max_it, 0, 1.0), "decode_host")

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()                      # returns when the outputs are on the host
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        te = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        assert torch.equal(h_it.to(dev), iters), "host-path iteration counts differ from the device path"
        out["e2e_app_llr"] = {"value": world * F * K * n_e2e / float(te.item()), "unit": "info bits/s",
                      "h2d_bytes_per_step": F * N * 4, "d2h_bytes_per_step": F * N * 4 + F * 4,
                      "ms_per_step": 1e3 * float(te.item()) / n_e2e,
                      "api": "lsb_ldpc_decode_host_f32 (C-ABI, pinned host LLR in, APP + iteration counts out)"}

        # same call returning hard decisions (what a BER Monte-Carlo consumes) instead of APP LLRs:
        # 4x fewer D2H bytes
        h_hard = torch.empty((F, N), dtype=torch.uint8).pin_memory()

        def e2e_hard_step():
            _lib.check(L.lsb_ldpc_decode_host_f32(g, h_llr.data_ptr(), None, h_hard.data_ptr(), h_it.data_ptr(), F,
                                                  args.max_it, 0, 1.0), "decode_host")

        e2e_hard_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_hard_step()
        th = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        assert torch.equal(h_hard.to(dev) != 0, app < 0), "host-path decisions differ from the device path"
        # headline e2e: the step's result is the decoded bits (what ldpc_awgn.py keeps: app < 0) + iteration
        # counts; the APP-LLR variant above returns the reference's raw decode() output and is host-bandwidth
        # bound once 8 ranks each write 510 MB per step into host memory
        out["e2e_u8_decisions"] = {"value": world * F * K * n_e2e / float(th.item()), "unit": "info bits/s",
                      "h2d_bytes_per_step": F * N * 4, "d2h_bytes_per_step": F * N + F * 4,
                      "ms_per_step": 1e3 * float(th.item()) / n_e2e,
                      "api": "lsb_ldpc_decode_host_f32 (C-ABI, pinned host LLR in, uint8 decisions + iteration counts out)"}

        # the same result as packed bits (32 decisions per word): 1/8 of the decision bytes back to the host
        pitch = (N + 31) // 32
        h_bits = torch.empty((F, pitch), dtype=torch.int32).pin_memory()

        def e2e_bits_step():
            _lib.check(L.lsb_ldpc_decode_host_bits_f32(g, h_llr.data_ptr(), h_bits.data_ptr(), h_it.data_ptr(), F,
                                                       args.max_it, 0, 1.0), "decode_host_bits")

        e2e_bits_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_bits_step()
        tb = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        shifts = torch.arange(32, device=dev, dtype=torch.int32)
        unpacked = ((h_bits.to(dev)[:, :, None] >> shifts) & 1).reshape(F, -1)[:, :N]
        assert torch.equal(unpacked != 0, app < 0), "bit-packed host-path decisions differ from the device path"
        assert torch.equal(h_it.to(dev), iters), "host-path iteration counts differ from the device path"
        del unpacked
        out["e2e"] = {"value": world * F * K * n_e2e / float(tb.item()), "unit": "info bits/s",
                      "h2d_bytes_per_step": F * N * 4, "d2h_bytes_per_step": F * pitch * 4 + F * 4,
                      "ms_per_step": 1e3 * float(tb.item()) / n_e2e,
                      "api": "lsb_ldpc_decode_host_bits_f32 (C-ABI, pinned host LLR in, bit-packed decisions + iteration counts out)"}

    out["clocks"] = clocks.summary()
    # Secondary rooflines (what the kernel is really bound by; messages never leave shared memory).
    # MUFU: the exp-domain check node spends dc EX2 + 2 dc LG2 per check = 3 per edge and iteration (it was
    # 4 (3 dc - 6) per check in round 1, when this pipe was the bound), against 16 MUFU lanes per SM per clock.
    # Issue: warp instructions per frame-iteration counted by ncu on this kernel and workload
    # (profiles/r2_bp_edomain_v2_ncu_full.csv: smsp__inst_executed.sum / frame-iterations of that launch), against
    # 4 issue slots per SM per clock.
    mufu_per_frame_iter = 3 * int(sum(int(dc) for dc in code.cdeg))
    sm_mhz = out["clocks"].get("sm_mhz") or out["clocks"].get("sm_max_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    mufu_peak = 16.0 * n_sm * float(sm_mhz) * 1e6
    mufu_rate = mufu_per_frame_iter * frame_iters_rank / (kernel_ms_avg * 1e-3)
    out["roofline_mufu"] = {"bound": "mufu", "achieved": mufu_rate / 1e12, "peak": mufu_peak / 1e12, "unit": "T MUFU op/s",
                            "frac": mufu_rate / mufu_peak, "mufu_per_frame_iteration": mufu_per_frame_iter,
                            "note": "16 MUFU lanes/SM/clk at the SM clock sampled under load"}
    issue_peak = 4.0 * n_sm * float(sm_mhz) * 1e6
    issue_rate = WARP_INSTR_PER_FRAME_ITER * frame_iters_rank / (kernel_ms_avg * 1e-3)
    out["roofline_issue"] = {"bound": "issue", "achieved": issue_rate / 1e12, "peak": issue_peak / 1e12,
                             "unit": "T warp-instructions/s", "frac": issue_rate / issue_peak,
                             "warp_instructions_per_frame_iteration": WARP_INSTR_PER_FRAME_ITER,
                             "note": "the real bound of the round-2 kernel: ncu smsp__issue_active 82 % "
                                     "(profiles/r2_bp_edomain_v2_ncu_full.csv); instruction count from that capture"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(code, args, args.cpu_seconds)
    if not args.no_extras:
        # the other BASELINE.json workloads (AMP half of the metric: configs 1, 3, 4, 5), on EVERY rank with the
        # headline's timing rules (tools/workloads.py): >= 5 steps / 2 warm-ups, barriers, max over ranks, clocks
        ctx = workloads.Ctx(rank, local_rank, world)
        out["extra_workloads"] = {}
        k = max(5, min(args.steps, 5))
        plan = (
            ("sparc_amp_config1", lambda: workloads.amp_workload(ctx, "c1", 16384, k, 2)),
            # same decode with the early-stop atol at the fp32 noise floor instead of the reference's 2e-15
            ("sparc_amp_config1_atol_1e-7", lambda: workloads.amp_workload(ctx, "c1", 16384, k, 2, atol=1e-7, e2e=False)),
            # config 4 (SC wave, w = 65536): the explicit-design tensor-core path is the throughput path; the FFT
            # path (exact fp32 arithmetic, 1e-4 parity per iteration) beside it on a smaller batch
            ("sparc_amp_config4_sc_wave", lambda: workloads.amp_workload(ctx, "c4", 512, k, 2, precision="tc")),
            ("sparc_amp_config4_sc_wave_fft_fp32", lambda: workloads.amp_workload(ctx, "c4", 148, 2, 1, e2e=False)),
            ("dense_amp_config3_tcgen05", lambda: workloads.dense_workload(ctx, 1024, k, 2)),
            ("ldpc_sparc_config5", lambda: workloads.c5_workload(ctx, 512, k, 2)),
        )
        for key, fn in plan:
            try:
                out["extra_workloads"][key] = fn()
            except Exception as e:   # the headline line must still be printed
                out["extra_workloads"][key] = {"error": repr(e)[:300]}
                if world > 1:
                    raise
            torch.cuda.empty_cache()
        # CPU baselines of the SPARC configs: the numpy oracle (a port of the reference, kind "port") timed on the
        # host cores over a bounded sample -- the cpu_baseline leg is allowed to execute oracle/
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            for key, cmd in (("sparc_amp_config1", ["c1", "--frames", str(4 * cores), "--procs", str(cores)]),
                             ("dense_amp_config3_tcgen05", ["c3", "--frames", "4"]),
                             ("sparc_amp_config4_sc_wave", ["c4", "--frames", "1"]),
                             ("ldpc_sparc_config5", ["c5", "--frames", "1"])):
                try:
                    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "cpu_time.py")] + cmd,
                                       capture_output=True, text=True, timeout=180, cwd=ROOT)
                    if isinstance(out["extra_workloads"].get(key), dict):
                        out["extra_workloads"][key]["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])
                except Exception as e:
                    out["extra_workloads"][key]["cpu_baseline"] = {"error": repr(e)[:200]}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Command-line entry to the Monte-Carlo campaigns the reference runs from its scripts.

    python -m ldpc_sparc_b200.campaign ldpc --standard 802.11n --rate 1/2 --z 81 --results data/results.txt
        ldpc_jossy/py/ldpc_awgn.py:60-123 (sim): SNR stepping until MIN_ERRORS block errors per point,
        one tuple per point appended to the results file in the reference's format
    python -m ldpc_sparc_b200.campaign sparc --L 1024 --M 32 --R 1.5 --frames 100000 [--pa B RATIO] [--sc OMEGA LAMBDA]
        sparc_public/sparc_sim.py over many frames of one design: BER / SER / frame error rate
    python -m ldpc_sparc_b200.campaign ldpc_sparc --mults 4 --frames 10000 --npz Test_36.npz
        ldpc_sparc/performance_plots_general.py:96-139: BER of the four schemes over an SNR grid

Under ``torchrun --nproc-per-node N`` every rank takes its share of the frames on its own GPU and
the error counters are all-reduced (NCCL); rank 0 prints / writes the results.
"""
import argparse
import json
import os

import numpy as np


def _init_dist():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return int(os.environ.get("RANK", "0")), world


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m ldpc_sparc_b200.campaign", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="what", required=True)
    p = sub.add_parser("ldpc")
    p.add_argument("--standard", default="802.11n")
    p.add_argument("--rate", default="1/2")
    p.add_argument("--z", type=int, default=81)
    p.add_argument("--ptype", default="A")
    p.add_argument("--points", type=int, default=24)
    p.add_argument("--min-errors", type=int, default=100)
    p.add_argument("--max-blocks", type=int, default=400000)
    p.add_argument("--max-it", type=int, default=200)
    p.add_argument("--results", default=None, help="file to append the result tuples to (reference: data/results.txt)")
    p.add_argument("--seed", type=int, default=0)
    p = sub.add_parser("sparc")
    p.add_argument("--L", type=int, default=1024)
    p.add_argument("--M", type=int, default=32)
    p.add_argument("--R", type=float, default=1.5)
    p.add_argument("--P", type=float, default=15.0)
    p.add_argument("--awgn-var", type=float, default=1.0)
    p.add_argument("--t-max", type=int, default=25)
    p.add_argument("--pa", nargs=2, metavar=("B", "R_PA_ratio"), default=None)
    p.add_argument("--sc", nargs=2, metavar=("omega", "Lambda"), default=None)
    p.add_argument("--K", type=int, default=1, help="K-PSK modulation of the sections (needs --complex for K > 2)")
    p.add_argument("--complex", action="store_true")
    p.add_argument("--frames", type=int, default=16384)
    p.add_argument("--precision", default="fp32")
    p.add_argument("--seed", type=int, default=0)
    p = sub.add_parser("ldpc_sparc")
    p.add_argument("--standard", default="802.16")
    p.add_argument("--rate", default="1/2")
    p.add_argument("--z", type=int, default=150)
    p.add_argument("--mults", type=int, default=4)
    p.add_argument("--logM", type=int, default=1)
    p.add_argument("--P", type=float, default=19.44)
    p.add_argument("--t-max", type=int, default=25)
    p.add_argument("--snr", type=float, nargs=3, default=[2.0, 6.0, 10], metavar=("start", "stop", "num"))
    p.add_argument("--frames", type=int, default=1024, help="frames per SNR point and scheme")
    p.add_argument("--precision", default="tc")
    p.add_argument("--npz", default=None, help="output file (reference: performance_plot_arrays/Test_N.npz)")
    p.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    rank, world = _init_dist()

    if args.what == "ldpc":
        from . import montecarlo
        res = montecarlo.sim(args.standard, args.rate, args.z, args.ptype, N_MEASUREMENTS=args.points,
                             MIN_ERRORS=args.min_errors, MAX_BLOCKS=args.max_blocks, results_path=args.results,
                             max_itcount=args.max_it, seed=args.seed)
        if rank == 0:
            for r in res:
                print(r)
    elif args.what == "sparc":
        from . import sparc_sim
        cp = dict(P=args.P, R=args.R, L=args.L, M=args.M)
        if args.pa:
            cp.update(power_allocated=True, B=int(args.pa[0]), R_PA_ratio=float(args.pa[1]))
        if args.sc:
            cp.update(spatially_coupled=True, omega=int(args.sc[0]), Lambda=int(args.sc[1]))
        if args.K > 1:
            cp.update(modulated=True, K=args.K)
        if args.complex:
            cp.update(complex=True)
        res = sparc_sim.sparc_sim_batch(cp, dict(t_max=args.t_max), args.awgn_var, args.frames, seed=args.seed,
                                        precision=args.precision)
        if rank == 0:
            print(json.dumps(res))
    else:
        from . import sparc_sim_new
        from .param_calc import param_calc
        int_rate = {"1/2": 1 / 2, "2/3": 2 / 3, "3/4": 3 / 4, "5/6": 5 / 6}[args.rate]
        M = 2 ** args.logM
        overall, L_sparc, L_sl, lengths = param_calc(args.mults, args.logM, args.standard, args.rate, int_rate, args.z, 1)
        sp = {'P': args.P, 'R': overall, 'L': L_sparc, 'M': M}
        spl = {'P': args.P, 'R': 1, 'L': L_sl, 'M': M}
        lp = {'standard': args.standard, 'rate': args.rate, 'z': args.z, 'int_rate': int_rate, 'mults': args.mults}
        snrs = np.linspace(args.snr[0], args.snr[1], int(args.snr[2]))
        res = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': args.t_max}, snrs, args.frames,
                                              seed=args.seed, precision=args.precision, save_path=args.npz)
        if rank == 0:
            for name, row in zip(res["sims_labels"], res["ber_store_averages"]):
                print("%-20s %s" % (name, " ".join("%.3e" % v for v in row)))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

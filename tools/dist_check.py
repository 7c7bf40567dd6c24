#!/usr/bin/env python
"""Rank-count independence of the sharded Monte-Carlo drivers on real GPUs / NCCL:

    python tools/dist_check.py --write ref.json                      (1 GPU: writes the totals)
    torchrun --nproc-per-node 2 tools/dist_check.py --check ref.json  (N GPUs: must reproduce them)

Frames are assigned to ranks by chunk and every chunk is seeded by (seed, point, chunk index), so
the all-reduced counters must be identical for any number of ranks."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def totals():
    from ldpc_sparc_b200 import ldpc, montecarlo, sparc_sim, sparc_sim_new
    from ldpc_sparc_b200.param_calc import param_calc
    out = {}
    c = ldpc.code("802.11n", "1/2", 27)
    r = montecarlo.ldpc_awgn_sweep(c, [1.0, 2.0], 4096, max_itcount=30, seed=3, chunk=512)
    out["ldpc_sweep"] = {k: [int(v) for v in r[k]] for k in montecarlo.COUNTER_FIELDS}
    r = sparc_sim.sparc_sim_batch(dict(P=15.0, R=1.45, L=256, M=32), dict(t_max=25), 1.0, 1024, seed=4, chunk=128)
    out["sparc_batch"] = {k: r[k] for k in sparc_sim.BATCH_FIELDS}
    r = sparc_sim.sparc_sim_batch(dict(P=15.0, R=2.4, L=128, M=16, complex=True, modulated=True, K=4), dict(t_max=20),
                                  1.0, 512, seed=5, chunk=64)
    out["csparc_batch"] = {k: r[k] for k in sparc_sim.BATCH_FIELDS}
    overall, L_sparc, L_sl, lengths = param_calc(1, 1, '802.16', '1/2', 1 / 2, 150, 1)
    sp = {'P': 19.44, 'R': overall, 'L': L_sparc, 'M': 2}
    spl = {'P': 19.44, 'R': 1, 'L': L_sl, 'M': 2}
    lp = {'standard': '802.16', 'rate': '1/2', 'z': 150, 'int_rate': 1 / 2, 'mults': 1}
    r = sparc_sim_new.performance_sweep(sp, spl, lp, lengths, {'t_max': 25}, [2.0, 4.0], 256, chunk=32, seed=6,
                                        precision="fp64")
    out["ldpc_sparc_sweep"] = {"bit_errors": r["bit_errors"].astype(int).tolist(), "bits": r["bits"].astype(int).tolist(),
                               "max": r["ber_store_max"].tolist(), "min": r["ber_store_min"].tolist()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--write")
    ap.add_argument("--check")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    got = totals()
    if (not world > 1) or dist.get_rank() == 0:
        if args.write:
            json.dump(got, open(args.write, "w"), indent=1)
            print("wrote", args.write, json.dumps(got)[:300])
        if args.check:
            ref = json.load(open(args.check))
            assert got == ref, "totals depend on the rank count:\n%s\n%s" % (json.dumps(got), json.dumps(ref))
            print("rank-count independence OK on %d ranks: %s" % (world, json.dumps(got)[:200]))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

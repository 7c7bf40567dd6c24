"""CPU: frame sharding + counter reduction, world_size 2 over gloo (host logic of
the N > 1 path; the decoder is stubbed -- the CUDA decoder is covered by -m gpu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ldpc_sparc_b200 import ldpc, montecarlo


def test_shards_partition_all_chunks():
    for n_chunks in (1, 2, 7, 16, 33):
        for world in (1, 2, 3, 8):
            got = [c for r in range(world) for c in montecarlo.shard_chunks(n_chunks, r, world)]
            assert got == list(range(n_chunks))
    assert montecarlo.chunk_ranges(10000, 4096) == [(0, 4096), (4096, 4096), (8192, 1808)]
    seeds = {montecarlo.chunk_seed(1, p, c) for p in range(4) for c in range(64)}
    assert len(seeds) == 256


def _stub_decode(llr, max_it):
    """Uncoded hard decision (keeps the test on the CPU)."""
    return (llr < 0), torch.full((llr.shape[0],), 3, dtype=torch.int32)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    code = ldpc.code("802.11n", "1/2", 27)
    res = montecarlo.ldpc_awgn_sweep(code, [0.0, 3.0], 5000, max_itcount=10, seed=7, device=torch.device("cpu"),
                                     decode_fn=_stub_decode, chunk=512)
    if rank == 0:
        q.put({k: v.tolist() for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sweep_equals_single_process():
    code = ldpc.code("802.11n", "1/2", 27)
    single = montecarlo.ldpc_awgn_sweep(code, [0.0, 3.0], 5000, max_itcount=10, seed=7, device=torch.device("cpu"),
                                        decode_fn=_stub_decode, chunk=512)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    multi = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in montecarlo.COUNTER_FIELDS:
        assert multi[k] == single[k].tolist(), k          # identical totals, independent of world size
    assert single["frames"].tolist() == [5000, 5000]
    assert single["iterations"].tolist() == [15000, 15000]
    # uncoded BPSK bit error rate Q(sqrt(1/sigma2)) at 0 dB ~ 0.1587, at 3 dB ~ 0.0789
    ber = single["bit_errors"] / (single["frames"] * code.N)
    assert abs(ber[0] - 0.1587) < 0.003 and abs(ber[1] - 0.0789) < 0.003

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


def _sim_worker(rank, world, port, q, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    code = ldpc.code("802.11n", "1/2", 27)
    enc = lambda u: torch.from_numpy(code.encode(u.numpy()).astype(np.uint8))
    res = montecarlo.sim("802.11n", "1/2", 27, N_MEASUREMENTS=2, MIN_ERRORS=50, MAX_BLOCKS=2000, results_path=path,
                         batch=256, device=torch.device("cpu"), decode_fn=_stub_decode, encode_fn=enc)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_campaign_driver_two_ranks_gloo(tmp_path):
    """montecarlo.sim over 2 gloo ranks with stubbed encode/decode: the per-round all-gather
    gives every rank the same global frame order, rank 0 writes the reference's tuple lines."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    path = str(tmp_path / "results.txt")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sim_worker, args=(r, 2, port, q, path)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lines = open(path).read().strip().splitlines()
    assert len(lines) == 2 and eval(lines[1]) == res[1]
    assert "np." not in lines[0] and lines[0].startswith("('802.11n', '1/2', 27, ")     # plain literals, as data/results.txt
    for std, rate, z, snr, nblocks, nerr, nbits, nbiterr, nit in res:
        assert nerr == 50 and nbits == nblocks * 324 and nit == 3 * nblocks      # uncoded stub: every block errs
        assert nblocks == 50
    assert abs(res[1][3] - res[0][3] - np.sqrt(100.0 / 50)) < 1e-12

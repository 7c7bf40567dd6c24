"""Frame sharding and error-counter reduction for Monte-Carlo runs on 1..8 GPUs.

The reference's drivers loop over frames serially (ldpc_jossy/py/ldpc_awgn.py:88-107,
ldpc_sparc/performance_plots_general.py:101-118) and frames never interact, so the
path shards by independent units: one process per GPU, contiguous ranges of fixed-
size frame chunks per rank, design data replicated, NO data-path collective.  The
only exchange is one all-reduce (sum) of a small int64 counter tensor
``[n_points, 5]`` = (bit errors, frame errors, section errors, frames, iterations)
at the end of a sweep -- NCCL over NVLink when the process group is NCCL, gloo in
the CPU tests.  Noise is drawn per chunk from a generator seeded by
(seed, point, global chunk index), so totals do not depend on the number of ranks.
"""
import numpy as np

COUNTER_FIELDS = ("bit_errors", "frame_errors", "section_errors", "frames", "iterations")
CHUNK = 4096   # frames per chunk: the unit of sharding and of RNG seeding


def chunk_ranges(n_frames, chunk=CHUNK):
    """[(first_frame, n)] covering n_frames in order."""
    return [(f0, min(chunk, n_frames - f0)) for f0 in range(0, n_frames, chunk)]


def shard_chunks(n_chunks, rank, world):
    """Contiguous block of chunk indices owned by `rank`: chunk c -> rank floor(c*world/n_chunks)."""
    lo = (rank * n_chunks + world - 1) // world
    hi = ((rank + 1) * n_chunks + world - 1) // world
    return range(lo, hi)


def chunk_seed(seed, point, chunk_index):
    """64-bit seed of one chunk's noise stream (SplitMix64 of the triple): independent of world size."""
    x = (int(seed) * 0x9E3779B97F4A7C15 + int(point) * 0xBF58476D1CE4E5B9 + int(chunk_index) * 0x94D049BB133111EB) & (2 ** 64 - 1)
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & (2 ** 64 - 1)
    x ^= x >> 31
    return x & (2 ** 63 - 1)


def all_reduce_counters(counters):
    """Sum a [n_points, 5] int64 tensor over all ranks (no-op without a process group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def ldpc_awgn_sweep(code, snr_db_points, frames_per_point, max_itcount=200, seed=0, device=None, decode_fn=None,
                    n_codewords=256, chunk=CHUNK):
    """BER / BLER / iteration statistics of ``code`` over BPSK-AWGN at each Es/N0
    (dB), the batched equivalent of ldpc_awgn.py:60-123 with a fixed frame budget.

    Every rank decodes its share of the chunks on its own GPU and the counters are
    all-reduced once.  Returns a dict of numpy arrays keyed by COUNTER_FIELDS plus
    'snr_db'.  ``decode_fn(llr_tensor, max_it) -> (hard uint8/bool [F,N], iters [F])``
    defaults to the CUDA decoder ``code.decode_hard``; the CPU tests inject a stub.
    """
    import torch
    rank, world = _rank_world()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    if decode_fn is None:
        decode_fn = lambda llr, it: code.decode_hard(llr, it)
    rng = np.random.RandomState(seed)
    x = code.encode(rng.randint(0, 2, (n_codewords, code.K)))          # pool of random codewords (host encoder)
    sgn = torch.from_numpy(1.0 - 2.0 * x).float().to(device)
    truth = torch.from_numpy(x.astype(np.uint8)).to(device)
    ranges = chunk_ranges(frames_per_point, chunk)
    counters = torch.zeros((len(snr_db_points), len(COUNTER_FIELDS)), dtype=torch.int64, device=device)
    for p, snr_db in enumerate(snr_db_points):
        sigma2 = 10.0 ** (-float(snr_db) / 10.0)                       # ldpc_awgn.py:45-49
        for c in shard_chunks(len(ranges), rank, world):
            f0, nf = ranges[c]
            gen = torch.Generator(device=device).manual_seed(chunk_seed(seed, p, c))
            which = (torch.arange(f0, f0 + nf, device=device) % n_codewords)
            noise = torch.randn((nf, code.N), device=device, generator=gen)
            llr = (2.0 / sigma2) * (sgn[which] + (sigma2 ** 0.5) * noise)  # ch2llr, ldpc_awgn.py:51-53
            hard, iters = decode_fn(llr, max_itcount)
            err = hard.to(torch.uint8) != truth[which]
            counters[p, 0] += err.sum()
            counters[p, 1] += err.any(dim=1).sum()
            counters[p, 3] += nf
            counters[p, 4] += iters.long().sum()
    all_reduce_counters(counters)
    out = {name: counters[:, i].cpu().numpy() for i, name in enumerate(COUNTER_FIELDS)}
    out["snr_db"] = np.asarray(snr_db_points, dtype=float)
    return out

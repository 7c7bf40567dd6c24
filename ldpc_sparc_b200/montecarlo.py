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
    fused = decode_fn is None            # the CUDA path: noise, LLRs, decode and error counts in one kernel
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
        if fused:
            # lsb_ldpc_mc_dev_f32: frame f of the point = codeword f mod n_codewords + Philox noise keyed by
            # (seed of the point, f); ranks take contiguous chunk ranges, totals do not depend on the world size
            mc = torch.zeros(4, dtype=torch.int64, device=device)
            for c in shard_chunks(len(ranges), rank, world):
                f0, nf = ranges[c]
                code.mc_decode(truth, nf, sigma2, seed=chunk_seed(seed, p, 0), frame0=f0, max_itcount=max_itcount,
                               counters=mc)
            counters[p, 0] += mc[0]
            counters[p, 1] += mc[1]
            counters[p, 3] += mc[2]
            counters[p, 4] += mc[3]
            continue
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


RATE_VALUE = {"1/2": .5, "2/3": 0.6667, "3/4": 0.75, "5/6": 0.83333}   # ldpc_awgn.py:63-72


def sim(standard, rate, z, ptype='A', N_MEASUREMENTS=24, C_AWGN_OFFSET=1.0, P_STEP=100.0, MIN_ERRORS=100,
        MAX_BLOCKS=400000, results_path=None, max_itcount=200, seed=0, batch=16384, device=None, decode_fn=None,
        encode_fn=None):
    """Batched equivalent of the reference campaign driver ``ldpc_awgn.sim``
    (ldpc_jossy/py/ldpc_awgn.py:60-123): for every measurement point random information
    words are encoded, sent over BPSK/AWGN and decoded until MIN_ERRORS block errors are
    seen (or MAX_BLOCKS blocks), one result tuple
    ``(standard, rate, z, SNR, nblocks, nblockerrors, nblocks*K, nbiterrors, nit_total)``
    is appended to ``results_path`` (the reference's data/results.txt line format, :109-114)
    and the SNR is stepped by ``sqrt(P_STEP / nblocks)`` (:116).

    Frames are generated, decoded and counted on the GPU in batches; the counters are cut
    at exactly the frame that brings the MIN_ERRORS-th block error, as the serial loop
    does.  With a process group every rank contributes ``batch`` frames per round (rank r
    takes the r-th slice of the round, so the global frame order is defined) and the
    per-frame statistics of a round are all-gathered.  Returns the list of tuples."""
    import torch
    import torch.distributed as dist
    from .ldpc import code as _code
    if rate not in RATE_VALUE:
        raise NameError("Rate unsupported")
    rank, world = _rank_world()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    mycode = _code(standard, rate, z, ptype)
    K, N = mycode.K, mycode.N
    fused = decode_fn is None            # CUDA path: channel, LLRs, decode and per-frame error counts in one kernel
    if decode_fn is None:
        decode_fn = lambda llr, it: mycode.decode_hard(llr, it)
    if encode_fn is None:
        encode_fn = lambda u: mycode.encode_device(u)
    SNR = 10.0 * np.log10(np.power(2, RATE_VALUE[rate]) - 1.0) + C_AWGN_OFFSET    # :76
    res = []
    for datapoint in range(N_MEASUREMENTS):
        sigma2 = 1.0 / np.power(10, SNR / 10.0)
        nblocks = nblockerrors = nbiterrors = nit_total = 0
        rnd = 0
        while nblockerrors < MIN_ERRORS and nblocks < MAX_BLOCKS:
            gen = torch.Generator(device=device).manual_seed(chunk_seed(seed, datapoint, rnd * world + rank))
            u = torch.randint(0, 2, (batch, K), device=device, generator=gen, dtype=torch.uint8)
            x = encode_fn(u)
            if fused:       # one codeword per frame; noise keyed by (seed of the round, global frame index)
                biterr, iters = mycode.mc_decode(x, batch, sigma2, seed=chunk_seed(seed, datapoint, rnd),
                                                 frame0=rank * batch, max_itcount=max_itcount)
            else:
                y = (1.0 - 2.0 * x.float()) + (sigma2 ** 0.5) * torch.randn((batch, N), device=device, generator=gen)
                hard, iters = decode_fn((2.0 / sigma2) * y, max_itcount)
                biterr = (hard.to(torch.uint8) != x).sum(dim=1).to(torch.int32)      # over all N code bits (:99)
            stats = torch.stack([biterr, iters.to(torch.int32)])                      # [2, batch]
            if world > 1:
                parts = [torch.empty_like(stats) for _ in range(world)]
                dist.all_gather(parts, stats.contiguous())
                stats = torch.cat(parts, dim=1)                                       # global frame order of the round
            be, it = stats[0].cpu().numpy().astype(np.int64), stats[1].cpu().numpy().astype(np.int64)
            room = MAX_BLOCKS - nblocks
            cum = np.cumsum(be > 0) + nblockerrors
            hit = np.flatnonzero(cum >= MIN_ERRORS)
            take = min(len(be), room, (hit[0] + 1) if len(hit) else len(be))
            nblocks += take
            nblockerrors += int((be[:take] > 0).sum())
            nbiterrors += int(be[:take].sum())
            nit_total += int(it[:take].sum())
            rnd += 1
        output = (standard, rate, z, float(SNR), int(nblocks), int(nblockerrors), int(nblocks) * int(K), int(nbiterrors),
                  int(nit_total))                      # plain ints: the line format of data/results.txt
        res.append(output)
        if results_path is not None and rank == 0:
            with open(results_path, 'a') as f:
                f.write(str(output))
                f.write("\n")
        SNR += np.sqrt(P_STEP / nblocks)
    return res

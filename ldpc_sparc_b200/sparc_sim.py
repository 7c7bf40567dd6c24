"""End-to-end SPARC simulations: host mirror of sparc_public/sparc_sim.py plus the
batched Monte-Carlo form that keeps frame generation, decoding and error counting on
the GPU.

``sparc_sim`` / ``calc_ber`` / ``calc_ser`` / ``awgn_channel`` keep the reference's names,
arguments and return values (sparc_sim.py:8-98,179-204).  ``sparc_sim_batch`` runs the same
experiment for ``n_frames`` frames that share one design: messages are drawn on the device,
``x = A beta0`` goes through the same DCT operator kernel the decoder uses, noise is added
on the device, the AMP kernel decodes, and only a handful of int64 counters leave the GPU
(all-reduced when a process group is up -- frames are sharded by chunk, no other exchange).
"""
import numpy as np

from . import montecarlo, sparc


# ------------------------------------------------------------------ reference API
def awgn_channel(input_array, awgn_var, rand_seed):
    """x + N(0, awgn_var), real or circularly-symmetric complex (sparc_sim.py:179-204)."""
    assert input_array.ndim == 1, 'input array must be one-dimensional'
    assert awgn_var >= 0
    rng = np.random.RandomState(rand_seed)
    n = input_array.size
    if input_array.dtype == float:
        return input_array + np.sqrt(awgn_var) * rng.randn(n)
    if input_array.dtype == complex:            # awgn_var / 2 per dimension
        return input_array + np.sqrt(awgn_var / 2) * (rng.randn(n) + 1j * rng.randn(n))
    raise Exception("Unknown input type '{}'".format(input_array.dtype))


def calc_ber(true_bin_array, est_bin_array):
    """Bit error rate of two bool arrays (sparc_sim.py:62-70)."""
    assert true_bin_array.dtype == 'bool'
    assert est_bin_array.dtype == 'bool'
    k = true_bin_array.size
    assert k == est_bin_array.size
    return np.count_nonzero(np.bitwise_xor(true_bin_array, est_bin_array)) / k


def calc_ser(beta0, beta, L):
    """Section error rate, locations and count (sparc_sim.py:72-98)."""
    assert beta.size == beta0.size, 'beta and beta0 are of different size'
    assert beta.dtype == beta0.dtype, 'beta and beta0 are of different type'
    assert type(L) == int and L > 0
    assert beta.size % L == 0
    wrong = (beta.reshape(L, -1) != beta0.reshape(L, -1)).any(axis=1)
    num_of_sec_errs = int(np.count_nonzero(wrong))
    return num_of_sec_errs / L, np.flatnonzero(wrong), num_of_sec_errs


def calc_ler_ver(beta0, beta, L, K):
    """Location / value / section error rates of a modulated SPARC (sparc_sim.py:100-175).
    Returns ``(ler, ver, ser), (loc_of_loc_errs, loc_of_val_errs, loc_of_sec_errs),
    (num_of_loc_errs, num_of_val_errs, num_of_sec_errs)``."""
    assert beta.size == beta0.size, 'beta and beta0 are of different size'
    assert beta.dtype == beta0.dtype, 'beta and beta0 are of different type'
    assert type(L) == int and L > 0
    assert beta.size % L == 0
    picked = []
    for vec in (beta0, beta):
        rows, cols = np.nonzero(vec.reshape(L, -1))
        assert np.array_equal(rows, np.arange(L))        # exactly 1 nonzero in each row
        picked.append((cols, vec.reshape(L, -1)[rows, cols]))
    loc_err = picked[0][0] != picked[1][0]
    val_err = picked[0][1] != picked[1][1]
    errs = (loc_err, val_err, loc_err | val_err)
    nums = tuple(int(np.count_nonzero(e)) for e in errs)
    return tuple(k / L for k in nums), tuple(np.flatnonzero(e) for e in errs), nums


def sparc_sim(code_params, decode_params, awgn_var, rand_seed=None, precision=None):
    """Encode, transmit and decode one frame (sparc_sim.py:8-58).  Returns the results dict
    ``ber, cer, t_final, nmse, detect, ser, loc_of_sec_errs, num_of_sec_errs`` (plus the
    location / value error entries ``ler, ver, ...`` for a modulated code).
    ``precision`` ('fp32' default / 'fp64') selects the decoder kernel, as in sparc_decode."""
    bits_i, beta0, x, Ab, Az = sparc.sparc_encode(code_params, awgn_var, rand_seed)
    y = awgn_channel(x, awgn_var, rand_seed)
    bits_o, beta, T, nmse, expect = sparc.sparc_decode(y, code_params, decode_params, awgn_var, rand_seed, beta0, Ab, Az,
                                                       precision=precision)
    ber = calc_ber(bits_i, bits_o)
    results = {'ber': ber, 'cer': 1.0 * (ber > 0), 't_final': T, 'nmse': nmse,
               'detect': 1.0 * (not (ber > 0) ^ bool(expect))}
    if not code_params['modulated']:
        ser, loc_of_sec_errs, num_of_sec_errs = calc_ser(beta0, beta, code_params['L'])
        results.update({'ser': ser, 'loc_of_sec_errs': loc_of_sec_errs, 'num_of_sec_errs': num_of_sec_errs})
    else:
        rates, locs, nums = calc_ler_ver(beta0, beta, code_params['L'], code_params['K'])
        results.update({'ser': rates[2], 'ler': rates[0], 'ver': rates[1],
                        'loc_of_sec_errs': locs[2], 'loc_of_loc_errs': locs[0], 'loc_of_val_errs': locs[1],
                        'num_of_sec_errs': nums[2], 'num_of_loc_errs': nums[0], 'num_of_val_errs': nums[1]})
    return results


# ------------------------------------------------------------------ batched Monte-Carlo
BATCH_FIELDS = ("bit_errors", "frame_errors", "section_errors", "frames", "iterations", "detected")


def sparc_sim_batch(code_params, decode_params, awgn_var, n_frames, rand_seed=(1, 2), seed=0, chunk=2048,
                    precision="fp32", device=None, n_designs=1):
    """``n_frames`` independent frames over a pool of ``n_designs`` designs.

    The reference draws a NEW design for every trial (sparc_sim.py:15-24 calls sparc_encode with the trial's
    ``rand_seed``, sparc.py:48,752), so its error rates average over the design ensemble as well as over messages
    and noise.  Here design ``d`` is drawn from ``rand_seed + d`` (element-wise) and block ``c`` of ``chunk`` frames
    runs on design ``c mod n_designs``: ``n_designs = 1`` is the fixed-design experiment (one design for all frames,
    the throughput configuration), larger pools approach the reference's ensemble average at the cost of
    ``n_designs`` sets of index tables (568 KB each at BASELINE config 1) and one kernel launch per block.

    Every rank handles its share of the ``chunk``-frame blocks (montecarlo.shard_chunks);
    the message and noise streams of a block are seeded by (seed, block index), so the totals
    do not depend on the number of ranks.  Returns a dict with the rates ``ber, cer, ser,
    detect``, the mean ``t_final`` and the raw int64 counters (BATCH_FIELDS)."""
    import torch
    sparc.check_decode_params(decode_params)
    code_params = dict(code_params)
    sparc.check_code_params(code_params)
    sparc._require_real_unmodulated(code_params)
    R, L, M = map(code_params.get, ['R', 'L', 'M'])
    K = sparc._modulation(code_params)
    cplx = code_params['complex']
    logM, logK = int(round(np.log2(M))), int(round(np.log2(K)))
    bit_len = L * (logM + logK)
    tmp = code_params.copy()
    tmp.update({'awgn_var': awgn_var})
    W = sparc.create_base_matrix(**tmp)
    n = int(round(bit_len / R))                                    # sparc.py:37-45
    if W.ndim == 2:
        n = int(round(n / W.shape[0])) * W.shape[0]
    code_params.update({'n': n, 'R_actual': bit_len / n})
    assert n_designs >= 1
    design_seeds = [[int(v) + d for v in rand_seed] for d in range(n_designs)]
    designs = {}

    def design(d):                       # built on first use: a rank only uploads the designs of its own blocks
        if d not in designs:
            designs[d] = sparc.sparc_transforms(W, L, M, n, design_seeds[d], code_params['complex'])
        return designs[d]

    rank, world = montecarlo._rank_world()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    rdt = torch.float64 if precision == "fp64" else torch.float32
    dt = rdt if not cplx else (torch.complex128 if precision == "fp64" else torch.complex64)
    counters = torch.zeros(len(BATCH_FIELDS), dtype=torch.int64, device=device)
    ranges = montecarlo.chunk_ranges(n_frames, chunk)
    shifts = torch.arange(logM + logK, device=device, dtype=torch.int32)
    constel = torch.as_tensor(np.asarray(sparc.psk_constel(K) if K > 1 else np.ones(1)), device=device).to(dt)

    def bit_pattern(sym):          # location bits, then the Gray code of the constellation index
        k = sym // M
        return ((sym % M) << logK) | (k ^ (k >> 1))

    for c in montecarlo.shard_chunks(len(ranges), rank, world):
        _, nf = ranges[c]
        Ab, Az = design(c % n_designs)
        rand_seed = design_seeds[c % n_designs]
        gen = torch.Generator(device=device).manual_seed(montecarlo.chunk_seed(seed, 0, c))
        idx0 = torch.randint(0, K * M, (nf, L), device=device, generator=gen, dtype=torch.int32)   # location + M * k
        beta0 = torch.zeros((nf, L * M), dtype=dt, device=device)
        beta0.view(nf, L, M).scatter_(2, (idx0 % M).long().unsqueeze(-1), constel[(idx0 // M).long()].unsqueeze(-1))
        x = Ab(beta0, precision)
        del beta0
        if cplx:                                                    # awgn_var / 2 per dimension
            noise = torch.view_as_complex(torch.randn((nf, n, 2), device=device, generator=gen, dtype=rdt))
            y = x + float(np.sqrt(awgn_var / 2)) * noise
        else:
            y = x + float(np.sqrt(awgn_var)) * torch.randn((nf, n), device=device, generator=gen, dtype=dt)
        _, idx, t_final, _, expect = sparc.sparc_decode(y, code_params, decode_params, awgn_var, list(rand_seed),
                                                        idx0, Ab, Az, precision=precision)
        diff = bit_pattern(idx) ^ bit_pattern(idx0)
        bit_err = ((diff.unsqueeze(-1) >> shifts) & 1).sum(dim=(1, 2))
        frame_err = bit_err > 0
        counters[0] += bit_err.sum()
        counters[1] += frame_err.sum()
        counters[2] += (diff != 0).sum()
        counters[3] += nf
        counters[4] += t_final.long().sum()
        counters[5] += (frame_err == expect).sum()                  # detect = not (err xor expect)
    montecarlo.all_reduce_counters(counters)
    cnt = {k: int(v) for k, v in zip(BATCH_FIELDS, counters.cpu().numpy())}
    frames = max(cnt["frames"], 1)
    out = {'ber': cnt["bit_errors"] / (frames * bit_len), 'cer': cnt["frame_errors"] / frames,
           'ser': cnt["section_errors"] / (frames * L), 't_final': cnt["iterations"] / frames,
           'detect': cnt["detected"] / frames, 'n': n, 'R_actual': bit_len / n, 'n_designs': n_designs}
    out.update(cnt)
    return out

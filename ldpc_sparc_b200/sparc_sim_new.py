"""Simulations of the LDPC-protected dense-design SPARC: host mirror of
sparc_sophie/sparc_sim_new.py (one frame per call) and the batched sweep that replaces
the loop of ldpc_sparc/performance_plots_general.py:96-139.

In the sweep all frames of a scheme share one design matrix (uploaded once), the user
bits are drawn, LDPC-encoded (lsb_ldpc_encode_dev), mapped to sections and multiplied by
A on the GPU, and the three decoders run on whole batches.  Ranks take disjoint chunks
of frames; bit-error counters are all-reduced once per sweep.
"""
import numpy as np

from . import montecarlo
from .ldpc import code
from .sparc_new import (_engine, bit_err_rate, create_design_matrix, integrated_decoder, naively_integrated_decoder,
                        sparc_ldpc_decode, sparc_ldpc_encode)


def awgn_channel(input_array, awgn_var, rand_seed):
    """x + N(0, awgn_var) (sparc_sim_new.py:212-224)."""
    assert input_array.ndim == 1, 'input array must be one-dimensional'
    assert awgn_var >= 0
    rng = np.random.RandomState(rand_seed)
    return input_array + np.sqrt(awgn_var) * rng.randn(input_array.size)


def _one_frame(decoder, sparc_params, ldpc_params, lengths, ldpc_bool, awgn_var, rand_seed):
    bits_i, total_bits, beta0, x, A = sparc_ldpc_encode(sparc_params, ldpc_params, lengths, ldpc_bool, rand_seed)
    y = awgn_channel(x, awgn_var, rand_seed)
    bits_o = decoder(y, A)
    return bits_i, bits_o, bit_err_rate(bits_i, bits_o)


def sparc_ldpc_sim(sparc_params, ldpc_params, lengths, ldpc_bool, decode_params, awgn_var, rand_seed=None):
    """AMP then BP (sparc_sim_new.py:12-23).  Returns ``bits_i, bits_o, ber``."""
    return _one_frame(lambda y, A: sparc_ldpc_decode(y, sparc_params, ldpc_params, decode_params, ldpc_bool, lengths, A),
                      sparc_params, ldpc_params, lengths, ldpc_bool, awgn_var, rand_seed)


def sparc_ldpc_naive_sim(sparc_params, ldpc_params, lengths, ldpc_bool, decode_params, awgn_var, rand_seed=None):
    """Alternating AMP / BP decoder (sparc_sim_new.py:52-63)."""
    return _one_frame(lambda y, A: naively_integrated_decoder(y, sparc_params, ldpc_params, decode_params, A),
                      sparc_params, ldpc_params, lengths, ldpc_bool, awgn_var, rand_seed)


def sparc_ldpc_integrated_sim(sparc_params, ldpc_params, lengths, ldpc_bool, decode_params, awgn_var, rand_seed=None):
    """AMP with the BP-aware denoiser (sparc_sim_new.py:141-152)."""
    return _one_frame(lambda y, A: integrated_decoder(y, sparc_params, ldpc_params, decode_params, A),
                      sparc_params, ldpc_params, lengths, ldpc_bool, awgn_var, rand_seed)


# ------------------------------------------------------------------ batched sweep
_sweep_engines = {}     # (L, M, n, design_seed, precision) -> device-resident design (generation takes seconds)


def _sweep_engine(L, M, n, design_seed, precision, keep=2):
    key = (L, M, n, design_seed, precision)
    if key not in _sweep_engines:
        if len(_sweep_engines) >= max(2, keep):
            _sweep_engines.clear()
        _sweep_engines[key] = _engine(create_design_matrix(L, M, n, design_seed), precision)
    return _sweep_engines[key]


SIMS = ('SPARC', 'SPARC+LDPC', 'Naive_decoder', 'Integrated_decoder')     # performance_plots_general.py:19


def _frames(eng, c, sparc_params, lengths, ldpc_bool, nf, gen):
    """Random user bits -> (user bits [nf, k] uint8, x [nf, n]) on the device (sparc_new.py:15-51)."""
    import torch
    P, L, M = sparc_params['P'], sparc_params['L'], sparc_params['M']
    logM = int(np.log2(M))
    dev = eng.dev
    if ldpc_bool:
        assert lengths['L_unprotected'] == 0, "only the fully protected scheme is built"
        mults = lengths['mults']
        user = torch.randint(0, 2, (nf, mults * c.K), device=dev, generator=gen, dtype=torch.uint8)
        total = c.encode_device(user.reshape(nf * mults, c.K)).reshape(nf, mults * c.N)
    else:
        user = torch.randint(0, 2, (nf, L * logM), device=dev, generator=gen, dtype=torch.uint8)
        total = user
    assert total.shape[1] == L * logM
    weights = (1 << torch.arange(logM - 1, -1, -1, device=dev)).long()          # first bit of a section is the MSB
    idx = (total.reshape(nf, L, logM).long() * weights).sum(dim=2)
    beta0 = torch.zeros((nf, L * M), dtype=eng.dt, device=dev)
    beta0.view(nf, L, M).scatter_(2, idx.unsqueeze(-1), float(np.sqrt(eng.n * P / L)))
    return user, eng.ab(beta0)


def performance_sweep(sparc_params, sparc_ldpc_params, ldpc_params, lengths, decode_params, snr_store,
                      frames_per_point, sims=SIMS, seed=0, design_seed=0, chunk=1024, precision="tc", save_path=None,
                      decode_params_integrated=None, n_designs=1):
    """BER of the four schemes of performance_plots_general.py:101-118 at every SNR of
    ``snr_store`` (linear, ``awgn_var = P / snr``), ``frames_per_point`` frames each.

    The reference draws a new design matrix for every run (sparc_new.py:48,1290 seeded per run by
    performance_plots_general.py:101-118), i.e. its BER curves average over the design ensemble.  ``n_designs`` > 1
    keeps a pool of that many device-resident designs (seeds ``design_seed + d``; 4 fp16 copies of A each, 3.3 GB at
    BASELINE config 5) and runs chunk ``c`` on design ``c mod n_designs``; 1 (default) is the fixed-design experiment.

    Returns a dict with ``snr_store``, ``ber_store_averages [len(sims), n_snr]``,
    ``ber_store_max`` / ``ber_store_min`` (extreme per-frame BERs, the reference's error
    bars) and the raw counters; with ``save_path`` rank 0 also writes the reference's
    ``.npz`` (same four array names, :137-138)."""
    import torch
    rank, world = montecarlo._rank_world()
    snr_store = np.asarray(snr_store, dtype=float)
    decode_params_integrated = decode_params_integrated or decode_params
    c = code(ldpc_params["standard"], ldpc_params["rate"], ldpc_params["z"])
    k_user = {False: sparc_params['L'] * int(np.log2(sparc_params['M'])), True: lengths['k_ldpc']}
    runs = {
        'SPARC': (sparc_params, False,
                  lambda y, A: sparc_ldpc_decode(y, sparc_params, ldpc_params, decode_params, False, lengths, A, precision)),
        'SPARC+LDPC': (sparc_ldpc_params, True,
                       lambda y, A: sparc_ldpc_decode(y, sparc_ldpc_params, ldpc_params, decode_params, True, lengths, A,
                                                      precision)),
        'Naive_decoder': (sparc_ldpc_params, True,
                          lambda y, A: naively_integrated_decoder(y, sparc_ldpc_params, ldpc_params, decode_params, A,
                                                                  precision)),
        'Integrated_decoder': (sparc_ldpc_params, True,
                               lambda y, A: integrated_decoder(y, sparc_ldpc_params, ldpc_params,
                                                               decode_params_integrated, A, precision)),
    }
    assert n_designs >= 1
    shapes = {}
    for ldpc_bool, sp in ((False, sparc_params), (True, sparc_ldpc_params)):
        if any(runs[s][1] == ldpc_bool for s in sims):
            L, M = sp['L'], sp['M']
            shapes[ldpc_bool] = (L, M, int(L * int(np.log2(M)) / sp['R']))

    def engine_of(ldpc_bool, ci):
        return _sweep_engine(*shapes[ldpc_bool], design_seed + ci % n_designs, precision, keep=len(shapes) * n_designs)

    dev = engine_of(next(iter(shapes)), 0).dev
    ranges = montecarlo.chunk_ranges(frames_per_point, chunk)
    mine = montecarlo.shard_chunks(len(ranges), rank, world)
    n_s, n_v = len(sims), len(snr_store)
    errs = torch.zeros((n_s, n_v, 2), dtype=torch.int64, device=dev)              # bit errors, bits
    ext = torch.zeros((n_s, n_v, 2), dtype=torch.float64, device=dev)             # max, -min of the per-frame BER
    ext[:, :, 1] = -1.0
    for v, snr in enumerate(snr_store):
        for ci in mine:
            _, nf = ranges[ci]
            made = {}
            for si, name in enumerate(sims):
                sp, ldpc_bool, dec = runs[name]
                eng = engine_of(ldpc_bool, ci)
                if ldpc_bool not in made:                                          # schemes on one design see the same frames
                    gen = torch.Generator(device=dev).manual_seed(montecarlo.chunk_seed(seed, v, ci))
                    user, x = _frames(eng, c, sp, lengths, ldpc_bool, nf, gen)
                    sigma = float(np.sqrt(sp['P'] / snr))
                    made[ldpc_bool] = (user, x + sigma * torch.randn(x.shape, device=dev, generator=gen, dtype=x.dtype))
                user, y = made[ldpc_bool]
                bits = torch.as_tensor(np.asarray(dec(y, eng)), device=dev).reshape(nf, -1)
                wrong = (bits.to(torch.uint8) != user).sum(dim=1)
                errs[si, v, 0] += wrong.sum()
                errs[si, v, 1] += nf * k_user[ldpc_bool]
                frame_ber = wrong.double() / k_user[ldpc_bool]
                ext[si, v, 0] = torch.maximum(ext[si, v, 0], frame_ber.max())
                ext[si, v, 1] = torch.maximum(ext[si, v, 1], -frame_ber.min())
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(errs, op=dist.ReduceOp.SUM)
        dist.all_reduce(ext, op=dist.ReduceOp.MAX)
    errs_h, ext_h = errs.cpu().numpy(), ext.cpu().numpy()
    out = {'snr_store': snr_store, 'ber_store_averages': errs_h[:, :, 0] / np.maximum(errs_h[:, :, 1], 1),
           'ber_store_max': ext_h[:, :, 0], 'ber_store_min': -ext_h[:, :, 1], 'bit_errors': errs_h[:, :, 0],
           'bits': errs_h[:, :, 1], 'sims_labels': list(sims)}
    if save_path is not None and rank == 0:
        np.savez(save_path, ber_store_averages=out['ber_store_averages'], ber_store_max=out['ber_store_max'],
                 ber_store_min=out['ber_store_min'], snr_store=snr_store)
    return out

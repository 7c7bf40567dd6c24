"""Property tests of the host-side helpers (CPU): bit maps, section symbols, frame sharding."""
import numpy as np
from hypothesis import given, settings, strategies as st

from ldpc_sparc_b200 import montecarlo, sparc

pow2 = st.sampled_from([2, 4, 8, 16, 32, 64, 512])
kpsk = st.sampled_from([1, 2, 4, 8, 16])


@settings(max_examples=60, deadline=None)
@given(M=pow2, K=kpsk, L=st.integers(1, 40), F=st.integers(0, 3), seed=st.integers(0, 2 ** 31 - 1))
def test_symbol_bit_maps_round_trip(M, K, L, F, seed):
    rng = np.random.RandomState(seed)
    sym = rng.randint(0, K * M, (F, L)) if F else rng.randint(0, K * M, L)
    bits = sparc.indices_2_bin_arr(sym, M, K)
    logM, logK = int(np.log2(M)), int(np.log2(K))
    assert bits.dtype == bool and bits.shape == sym.shape[:-1] + (L * (logM + logK),)
    assert np.array_equal(sparc.section_indices(bits, M, K), sym)
    if F == 0:                                                   # one frame: the reference's dense vectors
        vec = sparc.symbols_2_msg_vector(sym, M, K)
        assert vec.shape == (L * M,) and np.count_nonzero(vec) == L
        assert np.allclose(np.abs(vec[vec != 0]), 1.0)
        assert np.array_equal(sparc.msg_vector_2_symbols(vec, M, K), sym)
        assert np.array_equal(sparc.bin_arr_2_msg_vector(bits, M, K), vec)
        assert np.array_equal(sparc.msg_vector_2_bin_arr(vec, M, K), bits)
        assert np.array_equal(sparc.msg_vector_map_estimator(vec * 3.0, M, K), vec)     # MAP of a clean observation


@settings(max_examples=50, deadline=None)
@given(x=st.integers(0, 2 ** 40))
def test_gray_code_inverse(x):
    assert sparc.gray2bin(sparc.bin2gray(x)) == x
    y = x + 1
    assert bin(sparc.bin2gray(x) ^ sparc.bin2gray(y)).count("1") == 1          # neighbours differ in one bit


@settings(max_examples=100, deadline=None)
@given(n_frames=st.integers(0, 100000), chunk=st.integers(1, 5000), world=st.integers(1, 8))
def test_frame_sharding_partitions_the_frames(n_frames, chunk, world):
    ranges = montecarlo.chunk_ranges(n_frames, chunk)
    assert sum(n for _, n in ranges) == n_frames
    assert all(f0 == i * chunk and 0 < n <= chunk for i, (f0, n) in enumerate(ranges))
    owned = [list(montecarlo.shard_chunks(len(ranges), r, world)) for r in range(world)]
    flat = [c for o in owned for c in o]
    assert flat == list(range(len(ranges)))                       # disjoint, complete, contiguous per rank, in rank order
    sizes = [len(o) for o in owned]
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=50, deadline=None)
@given(seed=st.integers(0, 2 ** 31), point=st.integers(0, 100), chunk=st.integers(0, 10 ** 6))
def test_chunk_seed_is_a_63_bit_function_of_its_triple(seed, point, chunk):
    s = montecarlo.chunk_seed(seed, point, chunk)
    assert 0 <= s < 2 ** 63 and s == montecarlo.chunk_seed(seed, point, chunk)
    assert s != montecarlo.chunk_seed(seed, point, chunk + 1)

"""Edge cases on the GPU: empty and single-frame batches, frame counts that do not fill
a CTA wave or a GEMM tile, the smallest and largest section sizes, wrong shapes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ldpc_empty_single_and_ragged_batches():
    import torch
    from ldpc_sparc_b200 import ldpc
    c = ldpc.code("802.11n", "1/2", 27)
    rng = np.random.RandomState(0)
    x = c.encode(rng.randint(0, 2, (131, c.K)))
    llr = 2.0 * ((1 - 2.0 * x) + 0.7 * rng.randn(131, c.N)) / 0.49
    app0, it0 = c.decode(np.zeros((0, c.N)), 20)
    assert app0.shape == (0, c.N) and it0.shape == (0,)
    app_all, it_all = c.decode(llr, 20, precision="fp64")
    for F in (1, 2, 31, 131):                                   # every prefix decodes to the same values
        app, it = c.decode(llr[:F], 20, precision="fp64")
        assert np.array_equal(app, app_all[:F]) and np.array_equal(it, it_all[:F])
    a1, i1 = c.decode(llr[5], 20, precision="fp64")             # single-frame form returns (array, int)
    assert a1.shape == (c.N,) and isinstance(i1, int) and np.array_equal(a1, app_all[5])
    hard, it = c.decode_hard(torch.zeros((0, c.N), device="cuda"), 20)
    assert hard.shape == (0, c.N) and it.shape == (0,)
    with pytest.raises(NameError):
        c.decode(np.zeros(c.N + 1))
    with pytest.raises(NameError):
        c.decode(llr[0], 20, "no_such_decoder")
    assert c.encode_device(torch.zeros((0, c.K), dtype=torch.uint8, device="cuda")).shape == (0, c.N)


def test_ldpc_max_iterations_zero_and_one():
    from ldpc_sparc_b200 import ldpc
    import ldpc_oracle as lo
    c = ldpc.code("802.16", "5/6", 27)                          # check degree 20: the widest trellis bucket in use
    rng = np.random.RandomState(1)
    llr = rng.randn(9, c.N) * 3
    graph = lo.Graph(c.vdeg, c.cdeg, c.intrlv)
    for max_it in (0, 1, 3):
        app, it = c.decode(llr, max_it, precision="fp64")
        ref_app, ref_it = lo.decode_batch(graph, llr, max_it)
        assert np.array_equal(it, ref_it) and np.max(np.abs(app - ref_app)) < 1e-9 * max(1.0, np.max(np.abs(ref_app)))


def test_sparc_empty_and_ragged_batches():
    from ldpc_sparc_b200 import sparc
    cp = dict(P=15.0, R=1.3, L=64, M=8)
    bits, beta0, x, Ab, Az = sparc.sparc_encode(cp, 1.0, [1, 2])
    rng = np.random.RandomState(2)
    y = x[None, :] + rng.randn(37, x.size)
    full = sparc.sparc_decode(y, cp, dict(t_max=12), 1.0, [1, 2], None, Ab, Az, precision="fp64")
    for F in (1, 5, 37):
        part = sparc.sparc_decode(y[:F], cp, dict(t_max=12), 1.0, [1, 2], None, Ab, Az, precision="fp64")
        assert np.array_equal(part[0], full[0][:F]) and np.array_equal(part[2], full[2][:F])
    empty = sparc.sparc_decode(np.zeros((0, x.size)), cp, dict(t_max=12), 1.0, [1, 2], None, Ab, Az)
    assert empty[0].shape == (0, bits.size) and empty[1].shape == (0, 64)
    assert Ab(np.zeros((0, 64 * 8))).shape == (0, x.size) and Az(np.zeros((0, x.size))).shape == (0, 64 * 8)
    with pytest.raises(AssertionError):
        Ab(np.zeros(64 * 8 + 1))
    with pytest.raises(AssertionError):
        sparc.sparc_decode(np.zeros(x.size + 1), cp, dict(t_max=12), 1.0, [1, 2], None, Ab, Az)


@pytest.mark.parametrize("M", [2, 4, 1024])
def test_sparc_extreme_section_sizes(M):
    """M = 2 (sub-warp sections) and M = 1024 (32 entries per lane) against the oracle."""
    import sparc_oracle as so
    from ldpc_sparc_b200 import sparc
    L = 256 if M < 1024 else 16
    cp = dict(P=15.0, R=0.8 if M < 1024 else 1.2, L=L, M=M)
    seed = [4, 5]
    bits, beta0, x, Ab, Az = sparc.sparc_encode(dict(cp), 1.0, seed)
    rb, rbeta0, rx, design, n = so.sparc_encode(cp, 1.0, seed)
    assert np.array_equal(bits, rb) and np.max(np.abs(x - rx)) < 1e-10
    y = so.awgn_channel(rx, 1.0, seed)
    ref = so.sparc_decode(y, dict(so.normalise_code_params(cp), n=n), dict(t_max=15), 1.0, design, rbeta0)
    cpn = dict(cp)
    sparc.check_code_params(cpn)
    cpn.update(n=n)
    out = sparc.sparc_decode(y, cpn, dict(t_max=15), 1.0, seed, beta0, Ab, Az, precision="fp64")
    assert np.array_equal(out[0], ref[0]) and out[2] == ref[2]
    assert np.allclose(out[3], ref[3], rtol=1e-5, atol=1e-7)


def test_dense_ragged_frame_counts():
    """Frame counts around the 128-frame GEMM tile: tensor-core products against fp64."""
    import torch
    from ldpc_sparc_b200 import sparc_new
    L, M, n = 96, 8, 200                                         # n, LM multiples of 8 but not of the 128-wide tiles
    A = sparc_new.create_design_matrix(L, M, n, 3)
    rng = np.random.RandomState(3)
    for F in (1, 127, 129, 300):
        beta = rng.randn(F, L * M)
        z = rng.randn(F, n)
        for prec, tol in (("tc", 2e-4), ("fp32", 2e-5), ("fp64", 1e-12)):
            eng = sparc_new._engine(A, prec)
            ab = eng.ab(eng.to_dev(beta)).double().cpu().numpy()
            az = eng.az(eng.to_dev(z)).double().cpu().numpy()
            assert np.max(np.abs(ab - beta @ A.T)) <= tol * np.max(np.abs(beta @ A.T))
            assert np.max(np.abs(az - z @ A)) <= tol * np.max(np.abs(z @ A))
    eng = sparc_new._engine(A, "tc")
    assert eng.ab(torch.zeros((0, L * M), device="cuda")).shape == (0, n)


def test_state_evolution_small_sample_counts():
    import sparc_se_oracle as so
    from ldpc_sparc_b200 import sparc_se
    for mc in (1, 7, 257):
        u = np.random.RandomState(mc).randn(mc, 4)
        assert abs(sparc_se.sparc_se_E(0.7, 1, u) - so.se_expectation(0.7, 1, u)) < 1e-13

"""CPU: host-side logic of the ``code`` façade (no decode calls)."""
import numpy as np
import pytest

from ldpc_sparc_b200 import ldpc


def test_attributes_c2():
    c = ldpc.code("802.11n", "1/2", 81)
    assert (c.N, c.K, c.Nv, c.Nc, c.Nmsg) == (1944, 972, 1944, 972, 6966)
    assert sorted(set(c.cdeg)) == [7, 8] and sorted(set(c.vdeg)) == [2, 3, 4, 11]
    assert c.intrlv.dtype == np.int32
    assert sorted(c.intrlv.tolist()) == list(range(c.Nmsg))


def test_default_constructor_matches_reference_defaults():
    c = ldpc.code()
    assert (c.standard, c.rate, c.z, c.ptype) == ("802.11n", "1/2", 27, "A")
    assert c.N == 648


@pytest.mark.parametrize("args,msg", [
    (("802.11n", "1/2", 28), "802.11n invalid z"),
    (("802.11n", "7/8", 27), "802.11n invalid rate"),
    (("802.16", "7/8", 27), "802.16 invalid rate"),
    (("802.16", "2/3", 27, "C"), "802.16 type must be either A or B"),
    (("802.3", "1/2", 27), "IEEE standard unknown"),
])
def test_name_errors(args, msg):
    with pytest.raises(NameError, match=msg):
        ldpc.code(*args)


def test_decode_argument_errors_do_not_need_gpu():
    c = ldpc.code("802.11n", "1/2", 27)
    with pytest.raises(NameError, match="Channel inputs"):
        c.decode(np.zeros(c.N + 1))
    with pytest.raises(NameError, match="Decoder type"):
        c.decode(np.zeros(c.N), 10, "bogus")
    with pytest.raises(NameError, match="information word length"):
        c.encode(np.zeros(c.K + 1, dtype=int))


def test_batched_encode_is_linear():
    c = ldpc.code("802.16", "3/4", 54, "B")
    rng = np.random.RandomState(0)
    u, v = rng.randint(0, 2, (2, 8, c.K))
    assert np.array_equal((c.encode(u) + c.encode(v)) % 2, c.encode((u + v) % 2))
    assert np.array_equal(c.encode(u)[:, :c.K], u)   # systematic


def test_unpack_bits_layout():
    """Bit v % 32 of word v // 32, padding ignored (host helper of decode_bits; no GPU)."""
    from ldpc_sparc_b200 import ldpc
    c = ldpc.code("802.11n", "1/2", 27)                      # N = 648 = 20 words + 8 bits
    rng = np.random.RandomState(0)
    hard = rng.randint(0, 2, size=(5, c.N)).astype(np.uint8)
    pitch = (c.N + 31) // 32
    words = np.zeros((5, pitch), dtype=np.uint32)
    for v in range(c.N):
        words[:, v // 32] |= hard[:, v].astype(np.uint32) << np.uint32(v % 32)
    words[:, -1] |= np.uint32(0xFF000000)                    # garbage in the padding bits is dropped
    assert np.array_equal(c.unpack_bits(words), hard)

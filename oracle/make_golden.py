"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference under oracle/ref_shims.py) in the build container.  The fixtures
are committed; this script is the record of how they were made.

    python oracle/make_golden.py [ldpc] [sparc] [sophie] [--check-protographs]

Every fixture stores the inputs as well as the reference's outputs, so the tests
feed the very same arrays to the oracle and to the CUDA path (no dependence on
RNG stream stability across numpy versions).
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_shims  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def ldpc_channel(code, rng, snr_db, n_frames):
    """Random codewords over BPSK/AWGN, LLRs as ldpc_jossy/py/ldpc_awgn.py:45-56,88-95."""
    u = rng.randint(0, 2, (n_frames, code.K))
    x = np.stack([code.encode(ui) for ui in u])
    sigma2 = 1.0 / np.power(10, snr_db / 10.0)
    y = (1.0 - 2.0 * x) + np.sqrt(sigma2) * rng.randn(n_frames, code.N)
    return x.astype(np.uint8), 2.0 / sigma2 * y


def make_ldpc(ref):
    code = ref["ldpc"].code
    out = {}
    # (a) the BASELINE config-2 code at the three Es/N0 of SURVEY section 8(d)
    c = code("802.11n", "1/2", 81)
    rng = np.random.RandomState(20260118)
    for snr in (1.2, 1.6, 2.0):
        x, llr = ldpc_channel(c, rng, snr, 6)
        app = np.zeros_like(llr)
        its = np.zeros(len(llr), dtype=np.int32)
        for f in range(len(llr)):
            app[f], its[f] = c.decode(llr[f], 50)
        tag = "c2_snr%02d" % round(snr * 10)
        out[tag + "_x"], out[tag + "_llr"], out[tag + "_app"], out[tag + "_it"] = x, llr, app, its
    # (b) one frame of every (standard, rate) family at small z, sumprod2 and sumprod,
    #     both a noisy frame and the noiseless frame of test_ldpc.py:60
    fams = [("802.16", r, 27, p) for r, p in (("1/2", "A"), ("2/3", "A"), ("2/3", "B"), ("3/4", "A"), ("3/4", "B"), ("5/6", "A"))]
    fams += [("802.11n", r, 27, "A") for r in ("1/2", "2/3", "3/4", "5/6")]
    snr_for = {"1/2": 2.0, "2/3": 3.5, "3/4": 4.5, "5/6": 5.5}
    for i, fam in enumerate(fams):
        c = code(*fam)
        rng = np.random.RandomState(1000 + i)
        x, llr = ldpc_channel(c, rng, snr_for[fam[1]], 2)
        llr[1] = 10 * (0.5 - x[1])          # the reference test's noiseless input
        tag = "fam%d" % i
        out[tag + "_x"], out[tag + "_llr"] = x, llr
        for algo in ("sumprod2", "sumprod"):
            app = np.zeros_like(llr)
            its = np.zeros(2, dtype=np.int32)
            for f in range(2):
                app[f], its[f] = c.decode(llr[f], 40, algo)
            out[tag + "_app_" + algo], out[tag + "_it_" + algo] = app, its
    out["fam_names"] = np.array(["|".join(map(str, f)) for f in fams])
    # (c) Lxor / Lxfb known answers from the compiled reference
    L = ref["ldpc"].code("802.11n", "1/2", 27)
    rng = np.random.RandomState(7)
    pairs = np.concatenate([rng.randn(64, 2) * 6, [[0.0, 1.5], [-0.0, 2.0], [30.0, -31.0], [1e-9, -1e-9], [700.0, 700.0]]])
    out["lxor_in"] = pairs
    out["lxor_out"] = np.array([[L.Lxor(a, b, cf) for cf in (0, 1)] for a, b in pairs])
    vec = rng.randn(12, 9) * 4
    out["lxfb_in"] = vec
    res = [L.Lxfb(v, 1) for v in vec]
    out["lxfb_ret"] = np.array([r[0] for r in res])
    out["lxfb_out"] = np.stack([r[1] for r in res])
    np.savez_compressed(os.path.join(GOLD, "ldpc_reference.npz"), **out)
    print("ldpc_reference.npz:", {k: v.shape for k, v in out.items() if k.startswith("c2")})

    # (d) the reference's own graph-layout fixture ldpc_jossy/src/ldpc802.16.81.h
    hdr = open(os.path.join(ref_shims.REF_ROOT, "ldpc_jossy", "src", "ldpc802.16.81.h")).read()
    arrays = {}
    for name in ("intrlv", "vdeg", "cdeg"):
        body = re.search(r"long\s+%s\s*\[\s*\]\s*=\s*\{(.*?)\}" % name, hdr, re.S).group(1)
        arrays[name] = np.array([int(t) for t in re.findall(r"-?\d+", body)], dtype=np.int32)
    dims = {k: int(re.search(r"#define\s+%s\s+(\d+)" % k, hdr).group(1)) for k in ("Nv", "Nc", "Nmsg")}
    np.savez_compressed(os.path.join(GOLD, "ldpc_802_16_z81_graph.npz"), **arrays, **{k: np.array(v) for k, v in dims.items()})
    print("ldpc_802_16_z81_graph.npz:", dims)


def check_protographs(ref):
    from ldpc_sparc_b200 import ldpc as mine
    n = 0
    for std, zs in (("802.16", (3, 27, 54, 81, 150)), ("802.11n", (27, 54, 81))):
        for z in zs:
            for rate in ("1/2", "2/3", "3/4", "5/6"):
                for pt in ("A", "B"):
                    if pt == "B" and (std == "802.11n" or rate in ("1/2", "5/6")):
                        continue
                    r, c = ref["ldpc"].code(std, rate, z, pt), mine.code(std, rate, z, pt)
                    assert np.array_equal(r.proto, c.proto)
                    assert np.array_equal(r.intrlv, c.intrlv) and np.array_equal(r.vdeg, c.vdeg) and np.array_equal(r.cdeg, c.cdeg)
                    n += 1
    print("protographs and expanded graphs identical for %d codes" % n)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shims.load()
    args = sys.argv[1:] or ["ldpc", "sparc", "sophie"]
    if "--check-protographs" in args:
        check_protographs(ref)
    if "ldpc" in args:
        make_ldpc(ref)
    if "sparc" in args:
        from make_golden_sparc import make_sparc
        make_sparc(ref)
    if "sophie" in args:
        from make_golden_sparc import make_sophie
        make_sophie(ref)

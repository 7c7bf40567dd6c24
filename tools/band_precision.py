"""Per-iteration tau of the tensor-core AMP path (precision='tc') and of the fp32 FFT path against the reference's golden runs
(tests/golden/sparc_reference.npz): the numbers behind the precision paragraph of DESIGN.md section 5 (fp16 pairs, chunked
accumulation).  python tools/band_precision.py"""
import sys, os, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "oracle")); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
from ldpc_sparc_b200 import sparc
g = np.load(os.path.join(R, "tests/golden/sparc_reference.npz"))
meta = json.loads(str(g["meta_json"]))
def rel(a, b): return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1e-300)))
for name in ["c1_pa", "reg", "sc", "sc_pa_m16", "reg_m64", "big_reg", "c4_sc_wave"]:
    m = meta[name]
    for si in range(1 if name == "c4_sc_wave" else 2):
        pre = "%s_s%d_" % (name, si)
        cp, dp = dict(m["code_params"]), dict(m["decode_params"])
        seed = m["seeds"][si]
        bits_in, beta0, x, Ab, Az = sparc.sparc_encode(cp, m["awgn_var"], seed)
        sparc.check_decode_params(dp)
        out = {}
        for prec in ("tc", "fp32"):
            beta, T, nmse, psi, ex = sparc.sparc_amp(g[pre + "y"], cp, dp, m["awgn_var"], seed, beta0, Ab, Az, precision=prec, return_soft=True, return_trace=True)
            Tr = int(g[pre + "t_final"]); Tc = min(Tr, T)
            out[prec] = (T, rel(ex["tau"][:Tc], g[pre + "tau_it"][:Tc]), bool(np.array_equal(beta.reshape(-1, cp["M"]).argmax(1), g[pre + "idx"])))
        print(name, si, "T_ref", Tr, "| tc: T %d tau_err %.2e dec %s | fp32: T %d tau_err %.2e dec %s" % (out["tc"] + out["fp32"]))

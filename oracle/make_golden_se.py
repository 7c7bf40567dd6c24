"""State-evolution golden vectors from the UNMODIFIED reference (sparc_public/sparc_se.py,
imported under oracle/ref_shims.py).  For every case NumPy's global generator is seeded,
the reference's sparc_se is run, and psi / tau are stored together with the seed, so the
same call on the other side draws the same Monte-Carlo samples.

    python oracle/make_golden_se.py      (build container only; needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

SE_CASES = {
    "reg": (dict(P=15.0, R=1.3, M=32), 1.0, 25, 4000, 11),
    "reg_m512_low_snr": (dict(P=15.0, R=1.6, M=512), 1.0, 30, 1500, 12),
    "pa": (dict(P=15.0, R=1.5, M=32, power_allocated=True, B=32, R_PA_ratio=0.9), 1.0, 25, 3000, 13),
    "sc": (dict(P=15.0, R=1.4, M=32, spatially_coupled=True, omega=2, Lambda=8), 1.0, 25, 3000, 14),
    "sc_wave": (dict(P=15.0, R=1.5, M=512, spatially_coupled=True, omega=6, Lambda=32), 1.0, 40, 800, 15),
    "k2": (dict(P=15.0, R=1.3, M=32, modulated=True, K=2), 1.0, 25, 4000, 16),
    "k2_sc_pa": (dict(P=15.0, R=1.4, M=16, modulated=True, K=2, spatially_coupled=True, omega=3, Lambda=8,
                      power_allocated=True, B=2, R_PA_ratio=0.5), 1.0, 20, 3000, 17),
    "c_k1": (dict(P=15.0, R=2.6, M=32, complex=True), 1.0, 25, 3000, 18),
    "c_k4_pa": (dict(P=15.0, R=2.8, M=32, complex=True, modulated=True, K=4, power_allocated=True, B=8, R_PA_ratio=0.9),
                1.0, 25, 2000, 19),
    "c_k8_sc": (dict(P=15.0, R=3.0, M=16, complex=True, modulated=True, K=8, spatially_coupled=True, omega=3, Lambda=8),
                1.0, 25, 1500, 20),
}


def main():
    import ref_shims
    ref_shims.load()
    import sparc_se as ref_se          # sparc_public/sparc_se.py
    out, meta = {}, {}
    for name, (cp, var, t_max, mc, seed) in SE_CASES.items():
        np.random.seed(seed)
        psi, tau = ref_se.sparc_se(var, dict(cp), t_max, mc)
        out[name + "_psi"], out[name + "_tau"] = np.asarray(psi), np.asarray(tau)
        meta[name] = dict(code_params=cp, awgn_var=var, t_max=t_max, mc_samples=mc, seed=seed)
        print(name, "psi_final", np.round(np.atleast_1d(psi[-1])[:4], 6), "shape", np.shape(psi))
    out["meta_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "se_reference.npz"), **out)
    print("se_reference.npz: %.1f kB" % (os.path.getsize(os.path.join(GOLD, "se_reference.npz")) / 1e3))


if __name__ == "__main__":
    main()

import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np
import sophie_oracle as sph
from ldpc_sparc_b200 import sparc_new
for (L, M, n, F) in [(48, 2, 96, 130), (256, 4, 512, 64), (64, 64, 384, 5)]:
    P = 12.0
    rng = np.random.default_rng(L + M)
    A = rng.normal(0, 1 / np.sqrt(n), (n, L * M))
    sq = np.sqrt(n * P / L)
    sp, dp = dict(P=P, R=1.0, L=L, M=M), dict(t_max=12)
    beta0 = np.zeros((F, L * M))
    idx = rng.integers(0, M, (F, L))
    beta0[np.arange(F)[:, None], np.arange(L)[None, :] * M + idx] = sq
    ys = beta0 @ A.T + np.sqrt(P / 8.0) * rng.standard_normal((F, n))
    sparc_new.FUSED_TC_AMP = True
    b1, s1, t1 = sparc_new.sparc_amp(ys, sp, dp, A, precision="tc", return_trace=True)
    sparc_new.FUSED_TC_AMP = False
    b2, s2, t2 = sparc_new.sparc_amp(ys, sp, dp, A, precision="tc", return_trace=True)
    eb1 = eb2 = es1 = es2 = et1 = et2 = 0
    for f in range(min(F, 40)):
        tr = []
        rb, rs = sph.sparc_amp(ys[f], sp, dp, A, tr)
        rt = np.array([t["tau_sqr"] for t in tr])
        eb1 = max(eb1, np.max(np.abs(b1[f] - rb)) / sq); eb2 = max(eb2, np.max(np.abs(b2[f] - rb)) / sq)
        es1 = max(es1, np.max(np.abs(s1[f] - rs)) / sq); es2 = max(es2, np.max(np.abs(s2[f] - rs)) / sq)
        et1 = max(et1, np.max(np.abs(t1[f] - rt) / rt)); et2 = max(et2, np.max(np.abs(t2[f] - rt) / rt))
    print((L, M, n, F), "vs oracle: fused beta %.2e s %.2e tau %.2e | unfused beta %.2e s %.2e tau %.2e | fused-unfused beta %.2e s %.2e" %
          (eb1, es1, et1, eb2, es2, et2, np.max(np.abs(b1 - b2)) / sq, np.max(np.abs(s1 - s2)) / sq))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from ldpc_sparc_b200 import ldpc
c = ldpc.code("802.11n", "1/2", 81)
rng = np.random.RandomState(77)
x = c.encode(rng.randint(0, 2, (256, c.K)))
F, s2 = 16384, 10 ** (-1.2 / 10)
gen = torch.Generator(device="cuda").manual_seed(5)
sgn = torch.from_numpy(1.0 - 2.0 * x).float().cuda()
which = torch.arange(F, device="cuda") % 256
llr = (2 / s2) * (sgn[which] + (s2 ** 0.5) * torch.randn((F, c.N), device="cuda", generator=gen))
app32, it32 = c.decode(llr, 50, precision="fp32")
app64, it64 = c.decode(llr.double(), 50, precision="fp64")
conv = it64 < 50
rel = ((app32.double() - app64).abs() / app64.abs().clamp(min=1.0))
rel[~conv] = 0
fr = rel.max(1).values
order = torch.argsort(fr, descending=True)[:10]
for f in order.tolist():
    j = int(rel[f].argmax())
    print("frame", f, "it32", int(it32[f]), "it64", int(it64[f]), "maxrel %.3g" % float(fr[f]), "app64 %.5f app32 %.5f" % (float(app64[f, j]), float(app32[f, j])),
          "max|app64| %.1f" % float(app64[f].abs().max()), "n(rel>1e-3)", int((rel[f] > 1e-3).sum()))
print("frames with rel>1e-3:", int((fr > 1e-3).sum()), "of", int(conv.sum()))
print("quantiles of per-frame max rel:", [float(torch.quantile(fr[conv], q)) for q in (0.5, 0.9, 0.99, 0.999)])
print("max |app64| overall on converged:", float(app64[conv].abs().max()))

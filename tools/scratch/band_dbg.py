import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "oracle"))
import numpy as np, torch
import sparc_oracle as so
from ldpc_sparc_b200 import sparc
which = sys.argv[1]
W, L, M, n = so.sc_basic(np.array([2.0, 1.0]), 2, 3), 12, 8, 48
Ab, Az = sparc.sparc_transforms(W, L, M, n, [3, 4])
A = so.Design(W, L, M, n, [3, 4]).dense()
rng = np.random.RandomState(1)
x, y = rng.randn(5, L * M), rng.randn(5, n)
if which == "ab":
    ax = Ab(x, precision="tc"); print("ab err", np.max(np.abs(ax - x @ A.T)) / np.max(np.abs(x @ A.T)))
else:
    ay = Az(y, precision="tc"); print("az err", np.max(np.abs(ay - y @ A)) / np.max(np.abs(y @ A)))

"""Does a power-of-two scaling along K (V = (K_* C^-1)(W C)^T, C diagonal) save a digit plane?  CPU study.
   python tools/ozaki_colscale_study.py [N] [forest]"""
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nngp-src_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import nngp_oracle as orc  # noqa: E402
from nngp_b200 import synth  # noqa: E402
from ozaki_study import sliced_gemm_nt  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
if len(sys.argv) > 2:
    z = np.load(os.path.join(ROOT, "tests", "golden", "forest_xy.npz"))
    x, y, xt = z["x_train"][:n], z["y_train"][:n], z["x_test"][:384]
else:
    x, y, xt, _ = synth.make_problem(n, 384, 128)
fit = orc.Fit(x, y, depth=2)
mean, var = fit.predict(xt)
n = x.shape[0]
w = sla.solve_triangular(fit.c, np.eye(n), lower=True)
ks = orc.kernel_fn(xt, fit.x, 2)
kss = orc.final_diag(orc.layer0_diag(xt), 2)
r = np.abs(w) / np.max(np.abs(w), axis=1, keepdims=True)
low = r[np.tril_indices(n)]
print("fraction of W entries above rowmax/2^k:", {k: round(float((low > 2.0 ** -k).mean()), 4) for k in (1, 2, 4, 7, 10, 14)})
col = np.sqrt((w ** 2).sum(axis=0) / np.maximum(np.arange(n, 0, -1), 1))
c = np.exp2(-np.round(np.log2(col)))
for s in (5, 6, 7):
    e = []
    for a, b in ((ks, w), (ks / c[None, :], w * c[None, :])):
        v = sliced_gemm_nt(a, b, s)
        e.append(float(np.max(np.abs(kss - np.einsum("ij,ij->i", v, v) - var) / np.abs(var))))
    print(f"s={s}: plain {e[0]:.2e}   K-scaled {e[1]:.2e}")

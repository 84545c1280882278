"""Which plane pairs of the last kept group can be dropped?  W = L^-1 rows are dominated by their diagonal entry, so
plane 0 of W is nearly empty (digits 0 / +-1): the pair (A_{s-1}, W_0) weighs like a pair of the first DROPPED group.
   python tools/ozaki_pairskip_study.py [N] [forest]"""
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
from ozaki_study import split_rows  # noqa: E402

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


def product(s, skip):
    pa, ea = split_rows(ks, s)
    pb, eb = split_rows(w, s)
    acc = np.zeros((ks.shape[0], n))
    for g in range(s - 1, -1, -1):
        c = np.zeros_like(acc)
        for p in range(g + 1):
            if (p, g - p) in skip:
                continue
            c += pa[p] @ pb[g - p].T
        acc += np.ldexp(c, -7 * g)
    return np.ldexp(acc, (ea[:, None] - 6) + (eb[None, :] - 6))


for s in (6, 7):
    for name, skip in (("all pairs", set()), ("without (s-1, 0)", {(s - 1, 0)}), ("without (s-1,0),(s-2,0)", {(s - 1, 0), (s - 2, 0)}),
                       ("without (0, s-1)", {(0, s - 1)})):
        v = product(s, skip)
        e = float(np.max(np.abs(kss - np.einsum("ij,ij->i", v, v) - var) / np.abs(var)))
        print(f"s={s} {name:28s} {e:.2e}")

"""How many (256-row x 128-column) blocks of the digit planes of W = L^-1 (and of K_*) are entirely zero?
   python tools/plane_sparsity.py [N] [D] [depth]"""
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "nngp-src_b200"))
from oracle import nngp_oracle as orc  # noqa: E402
from nngp_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
xtr, ytr, xte, _ = synth.make_problem(n, 512, d)
fit = orc.Fit(xtr, ytr, depth=depth)
w = sla.solve_triangular(fit.c, np.eye(n), lower=True)
ks = orc.kernel_fn(xte, xtr, depth)


def planes(a, s=7):
    amax = np.max(np.abs(a), axis=1)
    _, e = np.frexp(amax)
    t = np.ldexp(a, (6 - e)[:, None])
    out = []
    for _ in range(s):
        q = np.rint(t)
        out.append(q)
        t = (t - q) * 128.0
    return out


for name, mat, tri in (("W", w, True), ("K*", ks, False)):
    pl = planes(mat)
    rows = mat.shape[0] // 256 * 256
    print(name, "row max / median |entry| (lower part):", float(np.median(np.max(np.abs(mat), axis=1) / np.median(np.abs(mat[mat != 0])))))
    for p, q in enumerate(pl[:4]):
        blk = np.abs(q[:rows, : n // 128 * 128]).reshape(rows // 256, 256, n // 128, 128).max(axis=(1, 3))
        if tri:
            mask = np.array([[kb * 128 < (rb + 1) * 256 for kb in range(n // 128)] for rb in range(rows // 256)])
            frac = float((blk[mask] == 0).mean())
            nz = float((np.abs(q[np.tril_indices(n)]) > 0).mean())
        else:
            frac = float((blk == 0).mean())
            nz = float((np.abs(q) > 0).mean())
        print(f"  plane {p}: all-zero 256x128 blocks {frac:.3f}   nonzero digits {nz:.3f}")

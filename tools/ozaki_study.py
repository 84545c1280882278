"""Accuracy study (CPU, numpy) for an int8-sliced evaluation of the variance GEMM  V = K_* · (L^-1)^T.

Each row of both operands is written as  2^e · 2^-6 · sum_p q_p 2^(-7p)  with int8 digits |q_p| <= 64 (one exponent
per row); the products of digit planes are exact integer GEMMs (int32 on the tensor cores; emulated here with FP64
matmuls of small integers, which are exact), planes with the same p+q share a scale, and only p+q < s is kept.
The script reports, for s = 4..8, the error of the posterior variance against the FP64 oracle.

    python tools/ozaki_study.py [n_train] [n_test] [forest]
"""
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import nngp_oracle as orc  # noqa: E402


def split_rows(a, s):
    """int digit planes q[p] (float arrays holding small integers) and per-row exponents e."""
    amax = np.max(np.abs(a), axis=1)
    _, e = np.frexp(amax)                     # amax = m 2^e, m in [0.5, 1)
    e = np.where(amax == 0, 0, e)
    t = np.ldexp(a, (6 - e)[:, None])         # |t| <= 64
    planes = []
    for _ in range(s):
        q = np.rint(t)
        planes.append(q)
        t = (t - q) * 128.0
    return planes, e


def sliced_gemm_nt(a, b, s):
    """a [T,K] · b[N,K]^T with s digit planes per operand, p+q < s."""
    pa, ea = split_rows(a, s)
    pb, eb = split_rows(b, s)
    acc = np.zeros((a.shape[0], b.shape[0]))
    for g in range(s - 1, -1, -1):            # small terms first
        c = np.zeros_like(acc)
        for p in range(g + 1):
            c += pa[p] @ pb[g - p].T
        assert np.max(np.abs(c)) < 2 ** 31, "int32 overflow"
        acc += np.ldexp(c, -7 * g)
    return np.ldexp(acc, (ea[:, None] - 6) + (eb[None, :] - 6))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    forest = len(sys.argv) > 3
    rng = np.random.default_rng(1)
    if forest:
        z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "forest_xy.npz"))
        x, y, xt = z["x_train"][:n], z["y_train"][:n], z["x_test"][:t]
        depth = 2
    else:
        d, depth = 128, 3
        x = rng.random((n, d))
        y = rng.normal(size=(n, 1)) * 3 + 8
        xt = np.vstack([rng.random((t - t // 4, d)), x[: t // 4] + 1e-3 * rng.normal(size=(t // 4, d))])
    fit = orc.Fit(x, y, depth=depth)
    mean, var = fit.predict(xt)
    ks = orc.kernel_fn(xt, fit.x, depth)
    w = sla.solve_triangular(fit.c, np.eye(fit.c.shape[0]), lower=True)
    kss = orc.final_diag(orc.layer0_diag(xt), depth)
    v64 = ks @ w.T
    var64 = kss - np.einsum("ij,ij->i", v64, v64)
    print(f"N={x.shape[0]} T={xt.shape[0]} cond~{np.linalg.cond(fit.c) ** 2:.2e} var range [{var.min():.3e}, {var.max():.3e}] kss~{kss.mean():.3f}")
    print(f"  explicit inverse, FP64 GEMM : max rel var err {np.max(np.abs(var64 - var) / np.abs(var)):.2e}  "
          f"std err {np.max(np.abs(np.sqrt(np.abs(var64)) - np.sqrt(var)) / np.sqrt(var)):.2e}")
    for s in (4, 5, 6, 7, 8):
        v = sliced_gemm_nt(ks, w, s)
        vs = kss - np.einsum("ij,ij->i", v, v)
        print(f"  s={s} ({s * (s + 1) // 2:2d} int8 GEMMs)      : max rel var err {np.max(np.abs(vs - var) / np.abs(var)):.2e}  "
              f"vs FP64-inverse path {np.max(np.abs(vs - var64) / np.abs(var)):.2e}  V max rel-to-rowmax {np.max(np.abs(v - v64)) / np.max(np.abs(v64)):.2e}")


if __name__ == "__main__":
    main()

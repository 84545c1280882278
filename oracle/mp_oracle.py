"""CPU ORACLE (high precision) -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED AGAINST THE REFERENCE.

50-digit mpmath restatement of SURVEY.md Appendix A (same algorithm as ``nngp_oracle.py``; see its
header for the reference call sites and why no reference-produced vectors exist).  Used once, by
``tests/golden/make_mp_golden.py``, to produce ``tests/golden/mp_small.npz``; the numpy oracle and the
CUDA path are both checked against those vectors.  Pure-Python loops: small cases only.
"""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 50


def _layers(depth, sigma_w, sigma_b):
    """[(sw2_l, sb2_l)] for Dense layer l: scalars apply to every layer, sequences give one value per layer."""
    sw = list(sigma_w) if hasattr(sigma_w, "__len__") else [sigma_w] * depth
    sb = list(sigma_b) if hasattr(sigma_b, "__len__") else [sigma_b] * depth
    assert len(sw) == depth and len(sb) == depth
    return [(mp.mpf(float(a)) ** 2, mp.mpf(float(b)) ** 2) for a, b in zip(sw, sb)]


def _diag0(x, sw2, sb2):
    d = len(x[0])
    return [sw2 * (mp.fsum(v * v for v in row) / d) + sb2 for row in x]


def kernel(x1, x2, depth, sigma_w, sigma_b, get="nngp"):
    x1 = [[mp.mpf(float(v)) for v in r] for r in x1]
    x2 = x1 if x2 is None else [[mp.mpf(float(v)) for v in r] for r in x2]
    lay = _layers(depth, sigma_w, sigma_b)
    sw2, sb2 = lay[0]
    d = len(x1[0])
    q1, q2 = _diag0(x1, sw2, sb2), _diag0(x2, sw2, sb2)
    out = []
    for i, a in enumerate(x1):
        row = []
        for j, b in enumerate(x2):
            sw2, sb2 = lay[0]
            k = sw2 * (mp.fsum(u * v for u, v in zip(a, b)) / d) + sb2
            ntk = k
            qa, qb = q1[i], q2[j]
            for layer in range(1, depth):
                sw2, sb2 = lay[layer]
                s2 = qa * qb - k * k
                s = mp.sqrt(s2) if s2 > 0 else mp.mpf(0)
                theta = mp.pi / 2 if (s == 0 and k == 0) else mp.atan2(s, k)
                dot_sigma = mp.mpf(1) / 2 - theta / (2 * mp.pi)
                k = sw2 * (s / (2 * mp.pi) + dot_sigma * k) + sb2
                ntk = k + sw2 * (ntk * dot_sigma)
                qa = sw2 * qa / 2 + sb2
                qb = sw2 * qb / 2 + sb2
            row.append(ntk if get == "ntk" else k)
        out.append(row)
    return out


def fit_predict(x, y, xt, depth, sigma_w, sigma_b, diag_reg, absolute=False):
    n = len(x)
    k = mp.matrix(kernel(x, None, depth, sigma_w, sigma_b))
    reg = mp.mpf(max(diag_reg, 0.0))
    lam = reg if absolute else reg * (mp.fsum(k[i, i] for i in range(n)) / n)
    for i in range(n):
        k[i, i] += lam
    c = mp.cholesky(k)
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    alpha = mp.cholesky_solve(k, yv)
    ks = mp.matrix(kernel(xt, x, depth, sigma_w, sigma_b))
    mean = ks * alpha
    ktt = kernel(xt, None, depth, sigma_w, sigma_b)
    var = []
    for i in range(len(xt)):
        v = mp.lu_solve(c, ks[i, :].T)  # C v = k_i (C lower triangular)
        var.append(ktt[i][i] - mp.fsum(e * e for e in v))
    return {"K": k, "lam": lam, "alpha": alpha, "Ks": ks, "mean": mean, "var": var}


def fit_predict_ntk(x, y, xt, depth, sigma_w, sigma_b, diag_reg, absolute=False):
    """NTK posterior at t = infinity, diagonal of the covariance."""
    n = len(x)
    th = mp.matrix(kernel(x, None, depth, sigma_w, sigma_b, "ntk"))
    kdd = mp.matrix(kernel(x, None, depth, sigma_w, sigma_b))
    reg = mp.mpf(max(diag_reg, 0.0))
    lam = reg if absolute else reg * (mp.fsum(th[i, i] for i in range(n)) / n)
    a = th.copy()
    for i in range(n):
        a[i, i] += lam
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    alpha = mp.cholesky_solve(a, yv)
    ths = mp.matrix(kernel(xt, x, depth, sigma_w, sigma_b, "ntk"))
    ks = mp.matrix(kernel(xt, x, depth, sigma_w, sigma_b))
    ktt = kernel(xt, None, depth, sigma_w, sigma_b)
    mean = ths * alpha
    var = []
    for i in range(len(xt)):
        w = mp.cholesky_solve(a, ths[i, :].T)
        quad = (w.T * kdd * w)[0]
        cross = mp.fsum(w[j] * ks[i, j] for j in range(n))
        var.append(ktt[i][i] + quad - 2 * cross)
    return {"Theta": th, "lam": lam, "alpha": alpha, "Thetas": ths, "mean": mean, "var": var}

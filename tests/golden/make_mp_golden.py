"""Generates tests/golden/mp_small.npz with the 50-digit mpmath oracle (oracle/mp_oracle.py).

Run from the repo root:  python tests/golden/make_mp_golden.py     (~1-2 minutes, CPU only)
Cases: seeded encoder-like inputs (nngp_b200.synth) small enough for pure-Python mpmath, covering the
reference configuration (depth 2, W_std 1, no bias, diag_reg 1e-3 relative), depth 3, the commented
reference variant W_std 1.5 / b_std 0.05 (active/active_train.py:44-49), depth 1, an absolute
regulariser, and degenerate rows (duplicates, an all-zero row, a scaled copy).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "nngp-src_b200"))
import mp_oracle as mpo  # noqa: E402
from nngp_b200 import synth  # noqa: E402

CASES = [
    # name, n, t, d, depth, sigma_w, sigma_b, diag_reg, absolute
    ("ref_d2", 40, 12, 20, 2, 1.0, 0.0, 1e-3, False),
    ("d3", 32, 10, 16, 3, 1.0, 0.0, 1e-3, False),
    ("sigma_variant", 32, 10, 16, 2, 1.5, 0.05, 1e-3, False),
    ("d1_linear", 24, 8, 10, 1, 1.0, 0.1, 1e-2, False),
    ("abs_reg", 24, 8, 10, 2, 1.0, 0.0, 5.0, True),
    ("degenerate", 24, 8, 8, 2, 1.0, 0.0, 1e-3, False),
]


def f64(m):
    return np.array([[float(m[i, j]) for j in range(m.cols)] for i in range(m.rows)])


def main():
    out = {}
    for name, n, t, d, depth, sw, sb, reg, absolute in CASES:
        xtr, ytr, xte, _ = synth.make_problem(n, t, d, seed_train=11, seed_test=12)
        if name == "degenerate":
            xtr[1] = xtr[0]            # duplicate training rows (s == 0, k > 0)
            xtr[2] = 0.0               # all-zero row (s == k == 0 -> theta = pi/2 branch)
            xtr[3] = 2.5 * xtr[0]      # scaled copy (cos theta = 1)
            xte[0] = xtr[0]            # test row equal to a training row
            xte[1] = 0.0
        r = mpo.fit_predict(xtr.tolist(), ytr.tolist(), xte.tolist(), depth, sw, sb, reg, absolute)
        kd = f64(r["K"])
        lam = float(r["lam"])
        kd[np.diag_indices(n)] -= lam   # store K_dd without the regulariser
        out[f"{name}/x_train"], out[f"{name}/y_train"], out[f"{name}/x_test"] = xtr, ytr, xte
        out[f"{name}/cfg"] = np.array([depth, sw, sb, reg, float(absolute)])
        out[f"{name}/K_dd"] = kd
        out[f"{name}/K_td"] = f64(r["Ks"])
        out[f"{name}/lam"] = np.array(lam)
        out[f"{name}/alpha"] = np.array([float(v) for v in r["alpha"]])
        out[f"{name}/mean"] = np.array([float(v) for v in r["mean"]])
        out[f"{name}/var"] = np.array([float(v) for v in r["var"]])
        print(name, "lam", lam, "mean[0]", out[f"{name}/mean"][0], "var[0]", out[f"{name}/var"][0])
    # NTK mode (train.py:254 `--kernel_type ntk`)
    for name, n, t, d, depth, sw, sb, reg in [("ntk_d2", 32, 10, 16, 2, 1.0, 0.0, 1e-3), ("ntk_d3_sigma", 28, 8, 12, 3, 1.5, 0.05, 1e-3)]:
        xtr, ytr, xte, _ = synth.make_problem(n, t, d, seed_train=21, seed_test=22)
        xtr[1] = xtr[0]; xte[0] = xtr[0]; xtr[2] = 0.0
        r = mpo.fit_predict_ntk(xtr.tolist(), ytr.tolist(), xte.tolist(), depth, sw, sb, reg)
        out[f"{name}/x_train"], out[f"{name}/y_train"], out[f"{name}/x_test"] = xtr, ytr, xte
        out[f"{name}/cfg"] = np.array([depth, sw, sb, reg, 0.0])
        out[f"{name}/Theta_dd"] = f64(r["Theta"])
        out[f"{name}/Theta_td"] = f64(r["Thetas"])
        out[f"{name}/lam"] = np.array(float(r["lam"]))
        out[f"{name}/alpha"] = np.array([float(v) for v in r["alpha"]])
        out[f"{name}/mean"] = np.array([float(v) for v in r["mean"]])
        out[f"{name}/var"] = np.array([float(v) for v in r["var"]])
        print(name, "lam", float(r["lam"]), "mean[0]", out[f"{name}/mean"][0], "var[0]", out[f"{name}/var"][0])
    np.savez_compressed(Path(__file__).with_name("mp_small.npz"), **out)


if __name__ == "__main__":
    main()

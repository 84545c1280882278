"""Generates tests/golden/mp_layers.npz with the 50-digit mpmath oracle: Dense layers that DIFFER in W_std / b_std
(neural-tangents allows it; the reference's own model never does) -- NNGP and NTK kernels and posteriors.

Run from the repo root:  python tests/golden/make_mp_layers_golden.py     (CPU only, < 1 minute)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "nngp-src_b200"))
import mp_oracle as mpo  # noqa: E402
from nngp_b200 import synth  # noqa: E402


def f64(m):
    return np.array([[float(m[i, j]) for j in range(m.cols)] for i in range(m.rows)])


def main():
    out = {}
    for name, n, t, d, sw, sb, reg in [("layers_d2", 24, 8, 10, (1.7, 0.6), (0.2, 0.4), 1e-3),
                                       ("layers_d3", 20, 6, 8, (1.0, 1.3, 0.8), (0.1, 0.0, 0.2), 1e-3)]:
        depth = len(sw)
        xtr, ytr, xte, _ = synth.make_problem(n, t, d, seed_train=31, seed_test=32)
        xtr[1] = xtr[0]; xtr[2] = 0.0; xte[0] = xtr[0]
        r = mpo.fit_predict(xtr.tolist(), ytr.tolist(), xte.tolist(), depth, sw, sb, reg)
        rn = mpo.fit_predict_ntk(xtr.tolist(), ytr.tolist(), xte.tolist(), depth, sw, sb, reg)
        kd = f64(r["K"])
        lam = float(r["lam"])
        kd[np.diag_indices(n)] -= lam
        out[f"{name}/x_train"], out[f"{name}/y_train"], out[f"{name}/x_test"] = xtr, ytr, xte
        out[f"{name}/sigma_w"], out[f"{name}/sigma_b"], out[f"{name}/diag_reg"] = np.array(sw), np.array(sb), np.array(reg)
        out[f"{name}/K_dd"], out[f"{name}/K_td"], out[f"{name}/lam"] = kd, f64(r["Ks"]), np.array(lam)
        out[f"{name}/mean"] = np.array([float(v) for v in r["mean"]])
        out[f"{name}/var"] = np.array([float(v) for v in r["var"]])
        out[f"{name}/Theta_dd"], out[f"{name}/Theta_td"] = f64(rn["Theta"]), f64(rn["Thetas"])
        out[f"{name}/ntk_lam"] = np.array(float(rn["lam"]))
        out[f"{name}/ntk_mean"] = np.array([float(v) for v in rn["mean"]])
        out[f"{name}/ntk_var"] = np.array([float(v) for v in rn["var"]])
        print(name, "lam", lam, "mean[0]", out[f"{name}/mean"][0], "ntk mean[0]", out[f"{name}/ntk_mean"][0])
    np.savez_compressed(Path(__file__).with_name("mp_layers.npz"), **out)


if __name__ == "__main__":
    main()

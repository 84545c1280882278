"""Print the parity numbers quoted in DESIGN.md / README.md: CUDA path vs the CPU oracle on the reference's forest
workload (C1, tests/golden/forest_xy.npz) and on a seeded synthetic C2 subsample.  GPU box only.
    python tests/checks/parity_report.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "nngp-src_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import nngp_oracle as oracle  # noqa: E402
from nngp_b200 import _lib, synth  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    out = {}
    f = np.load(ROOT / "tests" / "golden" / "forest_xy.npz")
    import time
    h = _lib.Handle()
    h.fit(f["x_train"], f["y_train"])
    mean, var = h.predict(f["x_test"])
    t0 = time.perf_counter(); h.fit(f["x_train"], f["y_train"]); t_fit = time.perf_counter() - t0
    t0 = time.perf_counter(); mean, var = h.predict(f["x_test"]); t_pred = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = oracle.Fit(f["x_train"], f["y_train"]); t_ofit = time.perf_counter() - t0
    t0 = time.perf_counter(); rm, rv = ref.predict(f["x_test"]); t_opred = time.perf_counter() - t0
    out["forest_c1"] = {"n_train": int(f["x_train"].shape[0]), "n_test": int(f["x_test"].shape[0]),
                        "mean_rel": rel(mean, rm), "var_rel": rel(var, rv), "std_rel": rel(np.sqrt(var), np.sqrt(rv)),
                        "alpha_rel": rel(h.get_state(x=False, l=False)["alpha"], ref.alpha),
                        "lambda": h.dims()[2], "lambda_oracle": float(ref.lam),
                        "gpu_fit_s": t_fit, "gpu_predict_s": t_pred, "oracle_fit_s": t_ofit, "oracle_predict_s": t_opred,
                        "host_cores": len(__import__("os").sched_getaffinity(0))}
    xtr, ytr, xte, _ = synth.make_problem(8192, 2048, 128)
    h.fit(xtr, ytr)
    mean, var = h.predict(xte)
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xte)
    out["synthetic_c2_sample"] = {"n_train": 8192, "n_test": 2048, "mean_rel": rel(mean, rm), "var_rel": rel(var, rv)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

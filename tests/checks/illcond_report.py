"""How the explicit 64 x 64 diagonal-block inverses of the solve -- and, in latency mode, the explicit inverse factor
L^-1 -- behave when K + lambda*I is badly conditioned: CUDA path vs the CPU oracle (LAPACK substitution) for
decreasing relative diag_reg.  GPU box only.
    python tests/checks/illcond_report.py"""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "nngp-src_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import nngp_oracle as oracle  # noqa: E402
from nngp_b200 import _lib, synth  # noqa: E402

xtr, ytr, xte, _ = synth.make_problem(3000, 1000, 32)
out = []
for reg in (1e-3, 1e-5, 1e-7, 1e-9):
    try:
        h = _lib.Handle(diag_reg=reg)
        h.fit(xtr, ytr)
        m, v = h.predict(xte)
        ref = oracle.Fit(xtr, ytr, diag_reg=reg)
        rm, rv = ref.predict(xte)
        k = oracle.kernel_fn(xtr)
        ev = np.linalg.eigvalsh(k + ref.lam * np.eye(3000))
        hl = _lib.Handle(diag_reg=reg, latency_mode=True)
        hl.fit(xtr, ytr)
        kss = oracle.final_diag(oracle.layer0_diag(xte))
        sliced = {}
        for sp in (7, 8, 9):          # variance_slices: the product on int8 digit planes (forced for these 1000 rows)
            hs = _lib.Handle(diag_reg=reg, variance_slices=sp)
            hs.fit(xtr, ytr)
            os.environ["NNGP_LATENCY_ROWS"] = "0"
            try:
                vs = hs.predict(xte)[1]
            finally:
                del os.environ["NNGP_LATENCY_ROWS"]
            sliced[str(sp)] = {"var_rel": float(np.max(np.abs(vs - rv) / np.abs(rv))), "var_rel_to_kss": float(np.max(np.abs(vs - rv)) / np.max(kss))}
            hs.close()
        vl = np.concatenate([hl.predict(xte[a:b])[1] for a, b in ((0, 1), (1, 8), (8, 300), (300, 1000))])   # GEMV, split-K, tiled
        out.append({"diag_reg": reg, "cond": float(ev[-1] / ev[0]), "mean_rel": float(np.max(np.abs(m - rm)) / np.max(np.abs(rm))),
                    "latency_mode_var_rel_to_kss": float(np.max(np.abs(vl - rv)) / np.max(kss)),
                    "latency_mode_var_rel": float(np.max(np.abs(vl - rv) / np.abs(rv))),
                    "var_rel_to_kss": float(np.max(np.abs(v - rv)) / np.max(oracle.final_diag(oracle.layer0_diag(xte)))),
                    "var_rel": float(np.max(np.abs(v - rv) / np.abs(rv))), "variance_slices": sliced, "min_var_over_kss": float(np.min(rv) / np.max(rv))})
    except Exception as e:  # noqa: BLE001
        out.append({"diag_reg": reg, "error": repr(e)[:200]})
print(json.dumps(out))

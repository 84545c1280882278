"""Prediction latency of medium batches: FP64 explicit-inverse path vs the int8 digit-plane path (variance_slices).
Host numpy buffers in, mean + variance out, wall clock around the C-ABI call.   python tools/sliced_sweep.py [N] [D] [s]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import _lib, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
S = int(sys.argv[3]) if len(sys.argv) > 3 else 7
xtr, ytr, xte, _ = synth.make_problem(N, 40000, D)
sizes = [128, 256, 384, 512, 768, 1024, 1536, 2048, 4096, 8192, 20000, 40000]


def sweep(h, reps=10):
    res = {}
    for t in sizes:
        x = np.ascontiguousarray(xte[:t])
        m, v = np.empty(t), np.empty(t)
        for _ in range(3):
            h.predict(x, mean_out=m, var_out=v)
        t0 = time.perf_counter()
        for _ in range(reps):
            h.predict(x, mean_out=m, var_out=v)
        res[str(t)] = round((time.perf_counter() - t0) / reps * 1e3, 4)
    return res


out = {"what": f"nngp_predict latency, N={N}, D={D}, depth 2, variance_slices={S}: wall ms per call (10 calls after 3 warm-ups)"}
h = _lib.Handle(stats_level=0, variance_slices=S)
h.fit(xtr, ytr)
os.environ["NNGP_LATENCY_ROWS"] = "1000000"
out["fp64_inverse_path_ms"] = sweep(h)
_m, v_ref = h.predict(xte[:4096])
os.environ["NNGP_LATENCY_ROWS"] = "0"
out["int8_planes_ms"] = sweep(h)
_m, v_s = h.predict(xte[:4096])
del os.environ["NNGP_LATENCY_ROWS"]
out["default_dispatch_ms"] = sweep(h)
out["var_max_rel_diff_4096_rows"] = float(np.max(np.abs(v_s - v_ref) / np.abs(v_ref)))
h.close()
hd = _lib.Handle(stats_level=0)
hd.fit(xtr, ytr)
out["substitution_path_ms"] = sweep(hd)
hd.close()
out["fp64_flop_floor_ms"] = {str(t): round(t * (float(N) * N + 2.0 * N * D) / 37.0e12 * 1e3, 4) for t in sizes}
print(json.dumps(out))

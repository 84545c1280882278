"""Where a single-query nngp_predict call spends its time (latency mode): per-stage device times (stats_level 1)
next to the wall clock of the call.   python tools/latency_probe.py [N] [D] [T]"""
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
T = int(sys.argv[3]) if len(sys.argv) > 3 else 1
xtr, ytr, xte, _ = synth.make_problem(N, max(T, 8), D)
for level in (0, 1):
    h = _lib.Handle(stats_level=level, latency_mode=True)
    h.fit(xtr, ytr)
    x = np.ascontiguousarray(xte[:T])
    m, v = np.empty(T), np.empty(T)
    for _ in range(5):
        h.predict(x, mean_out=m, var_out=v)
    h.stats_reset()
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        h.predict(x, mean_out=m, var_out=v)
    wall = (time.perf_counter() - t0) / reps * 1e3
    s = h.stats()
    print(f"stats_level {level}: wall {wall:.4f} ms/call; device stages per call (ms):",
          {k: round(s[k] / reps, 4) for k in ("pred_total_ms", "pred_gram_ms", "pred_trsm_ms", "h2d_ms", "d2h_ms")},
          "launches/call", s["kernel_launches"] / reps)
    h.close()

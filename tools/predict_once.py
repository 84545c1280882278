"""One fit at N = 8192 and a few predictions of T rows -- a small driver for profiling a given batch size.
    python tools/predict_once.py [T] [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import _lib, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
xtr, ytr, xte, _ = synth.make_problem(8192, T, 128)
h = _lib.Handle()
h.fit(xtr, ytr)
for _ in range(reps):
    mean, var = h.predict(xte)
print("T", T, "mean[0]", mean[0], "var[0]", var[0], "pred_total_ms", h.stats()["pred_total_ms"] / reps)

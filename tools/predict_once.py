"""One fit and a few predictions of T device-resident rows -- a small driver for profiling a given shape.
    python tools/predict_once.py [T] [reps] [N] [D] [depth] [variance_slices]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import _lib, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
D = int(sys.argv[4]) if len(sys.argv) > 4 else 128
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 2
slices = int(sys.argv[6]) if len(sys.argv) > 6 else 0
xtr = synth.encodings(N, D, 1)
ytr = synth.labels(xtr)
xte = torch.from_numpy(synth.encodings(T, D, 2)).cuda()
mean = torch.empty(T, dtype=torch.float64, device="cuda")
var = torch.empty(T, dtype=torch.float64, device="cuda")
h = _lib.Handle(depth=depth, stats_level=1, variance_slices=slices)
h.fit(xtr, ytr)
h.predict(xte, mean_out=mean, var_out=var)
h.stats_reset()
for _ in range(reps):
    h.predict(xte, mean_out=mean, var_out=var)
s = h.stats()
print("N", N, "D", D, "depth", depth, "T", T, "mean[0]", float(mean[0]), "var[0]", float(var[0]),
      "pred_total_ms", s["pred_total_ms"] / reps, "trsm_ms", s["pred_trsm_ms"] / reps, "gram_ms", s["pred_gram_ms"] / reps, "sliced_ms", s["sliced_ms"] / reps,
      "int8_tops", 2 * s["sliced_macs"] / max(s["sliced_ms"], 1e-9) / 1e9)

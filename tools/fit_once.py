"""One warm-up fit and `reps` timed fits at N rows (D features, depth) -- a small driver for ncu launch lists of the
fit path.   python tools/fit_once.py [N] [D] [depth] [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import _lib, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
x = synth.encodings(N, D, 1)
y = synth.labels(x)
h = _lib.Handle(depth=depth, stats_level=1)
h.fit(x, y)
h.stats_reset()
for _ in range(reps):
    h.fit(x, y)
s = h.stats()
print("N", N, "fit_total_ms", s["fit_total_ms"] / reps, "chol_ms", s["fit_chol_ms"] / reps, "gram_ms", s["fit_gram_ms"] / reps,
      "solve_ms", s["fit_solve_ms"] / reps, "launches", s["kernel_launches"] / reps)

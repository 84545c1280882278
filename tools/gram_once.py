"""Gram kernel driver: K(x1, x2) at (M, N, D, depth) on device buffers, `reps` launches (for ncu / timing).
    python tools/gram_once.py [M] [N] [D] [depth] [reps]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import _lib, synth  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
x1 = torch.from_numpy(synth.encodings(M, D, 2)).cuda()
x2 = torch.from_numpy(synth.encodings(N, D, 1)).cuda()
out = torch.empty((M, N), dtype=torch.float64, device="cuda")
h = _lib.Handle(depth=depth, stats_level=2)
h.kernel(x1, x2, out=out)
h.stats_reset()
for _ in range(reps):
    h.kernel(x1, x2, out=out)
s = h.stats()
peak = h.dmma_peak_tflops()
tf = s["gram_flops"] / s["gram_ms"] / 1e9
print(f"gram M={M} N={N} D={D} depth={depth}: {s['gram_ms'] / reps:.3f} ms/launch  {tf:.2f} TFLOP/s  frac {tf / peak:.3f}  "
      f"evals/s {s['gram_evals'] / s['gram_ms'] * 1e3:.3e}")

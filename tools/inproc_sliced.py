"""ONE handle on G GPUs (replicas + row split behind nngp_predict) with the int8 digit-plane variance path, at the C3
per-GPU shape.   python tools/inproc_sliced.py [G] [slices] [N] [T_per_gpu] [D] [depth]"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import _lib, synth  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 7
N = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
TG = int(sys.argv[4]) if len(sys.argv) > 4 else 131072
D = int(sys.argv[5]) if len(sys.argv) > 5 else 256
depth = int(sys.argv[6]) if len(sys.argv) > 6 else 3
xtr = synth.encodings(N, D, 1)
ytr = synth.labels(xtr)
T = G * TG
xte = torch.from_numpy(synth.encodings(T, D, 2)).pin_memory()
mean = torch.empty(T, dtype=torch.float64).pin_memory()
var = torch.empty(T, dtype=torch.float64).pin_memory()
out = {"gpus": G, "variance_slices": S, "n_train": N, "test_rows": T, "dim": D, "depth": depth}
res = {}
for name, g in (("one_gpu", 1), ("handle_on_%d_gpus" % G, G)):
    h = _lib.Handle(depth=depth, stats_level=1, n_gpus=g, variance_slices=S)
    h.fit(xtr, ytr)
    h.stats_reset()
    h.fit(xtr, ytr)
    st = h.stats()
    t_rows = TG if g == 1 else T
    x, m, v = xte.numpy()[:t_rows], mean.numpy()[:t_rows], var.numpy()[:t_rows]
    h.predict(x, mean_out=m, var_out=v)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        h.predict(x, mean_out=m, var_out=v)
    dt = (time.perf_counter() - t0) / reps
    res[name] = (m.copy(), v.copy())
    out[name] = {"queries_per_s": t_rows / dt, "predict_s": dt, "fit_total_ms": st["fit_total_ms"], "inverse_ms": st["inverse_ms"],
                 "replicate_ms": st["replicate_ms"], "replicate_gb_per_s": st["replicate_bytes"] / max(st["replicate_ms"], 1e-9) / 1e6}
    h.close()
    torch.cuda.empty_cache()
m1, v1 = res["one_gpu"]
mg, vg = res["handle_on_%d_gpus" % G]
out["first_gpu_share_bitwise_equal"] = bool(np.array_equal(m1, mg[:TG]) and np.array_equal(v1, vg[:TG]))
out["scaling"] = out["handle_on_%d_gpus" % G]["queries_per_s"] / out["one_gpu"]["queries_per_s"]
print(json.dumps(out))

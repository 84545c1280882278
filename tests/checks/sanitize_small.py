"""Small end-to-end pass of every kernel of the hot path, sized for compute-sanitizer (memcheck / racecheck /
synccheck): kernel_fn, fit (Gram, blocked Cholesky with look-ahead, solves), predict (fused and stepwise TRSM)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "nngp-src_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))
import nngp_oracle as oracle  # noqa: E402
from nngp_b200 import _lib, synth  # noqa: E402

n, t, d = int(os.environ.get("SAN_N", 700)), int(os.environ.get("SAN_T", 300)), 24
xtr, ytr, xte, _ = synth.make_problem(n, t, d)
h = _lib.Handle(depth=3, sigma_w=1.2, sigma_b=0.1)
k = h.kernel(xtr[:150], xte[:70])
h.fit(xtr, ytr)
mean, var = h.predict(xte)
ref = oracle.Fit(xtr, ytr, 3, 1.2, 0.1)
rm, rv = ref.predict(xte)
print("kernel", np.max(np.abs(k - oracle.kernel_fn(xtr[:150], xte[:70], 3, 1.2, 0.1))) / np.max(k))
print("mean", np.max(np.abs(mean - rm)) / np.max(np.abs(rm)), "var", np.max(np.abs(var - rv)) / np.max(rv))
print("lml", h.log_marginal_likelihood(), ref.log_marginal_likelihood())
st = h.get_state()
h2 = _lib.Handle(depth=3, sigma_w=1.2, sigma_b=0.1, max_block_bytes=128 * 704 * 8)
h2.set_state(st["x"], st["l"], st["alpha"], st["lambda"])
m2, v2 = h2.predict(xte)
print("blocked == unblocked", np.array_equal(m2, mean), np.array_equal(v2, var))
print("SANITIZE PASS DONE")

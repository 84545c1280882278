"""Two host threads, two handles on ONE GPU, both predicting through the int8 digit-plane path (cooperative launches
with a grid-wide re-alignment counter) at the same time: no deadlock, and every result bitwise the serial one.
    python tests/checks/sliced_soak.py [iterations]"""
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import _lib, synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
xtr, ytr, xte, _ = synth.make_problem(4096, 20000, 64)
hs = [_lib.Handle(variance_slices=7, stats_level=0) for _ in range(2)]
for h in hs:
    h.fit(xtr, ytr)
m0, v0 = hs[0].predict(xte)
bad = [0, 0]


def work(i):
    for _ in range(iters):
        m, v = hs[i].predict(xte)
        if not (np.array_equal(m, m0) and np.array_equal(v, v0)):
            bad[i] += 1


t0 = time.time()
ths = [threading.Thread(target=work, args=(i,)) for i in range(2)]
for t in ths:
    t.start()
for t in ths:
    t.join()
print(json.dumps({"concurrent_predictions": 2 * iters, "differ_from_serial": sum(bad), "seconds": round(time.time() - t0, 2),
                  "rows": int(xte.shape[0]), "n_train": 4096}))
sys.exit(1 if sum(bad) else 0)

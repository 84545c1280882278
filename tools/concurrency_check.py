"""Two handles fitting / predicting concurrently from two host threads (two CUDA streams of one process), compared
bit for bit with a serial reference.  Exposes any cross-stream interference between this library's kernels."""
import sys
import threading
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import _lib, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nthreads = int(sys.argv[3]) if len(sys.argv) > 3 else 2
xtr, ytr, xte, _ = synth.make_problem(n, 8192, 128)
ytr = np.ones_like(ytr)
ref_h = _lib.Handle()
ref_h.fit(xtr, ytr)
ref = (ref_h.get_state(x=False, l=False)["alpha"],) + tuple(ref_h.predict(xte))
bad = {"fit": 0, "predict": 0}
lock = threading.Lock()


def worker(tid):
    h = _lib.Handle()
    for r in range(reps):
        h.fit(xtr, ytr)
        a = h.get_state(x=False, l=False)["alpha"]
        m, v = h.predict(xte)
        with lock:
            if not np.array_equal(a, ref[0]):
                bad["fit"] += 1
                print(f"thread {tid} rep {r}: alpha differs, rel {np.max(np.abs(a - ref[0])) / np.max(np.abs(ref[0])):.3e}", flush=True)
            elif not (np.array_equal(m, ref[1]) and np.array_equal(v, ref[2])):
                bad["predict"] += 1
                print(f"thread {tid} rep {r}: prediction differs", flush=True)


ts = [threading.Thread(target=worker, args=(i,)) for i in range(nthreads)]
for t in ts:
    t.start()
for t in ts:
    t.join()
print(f"N={n} threads={nthreads} reps={reps}: mismatching fits {bad['fit']}, mismatching predictions {bad['predict']}")

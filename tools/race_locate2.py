"""Compare the partially factored matrix after k outer steps with / without look-ahead (debug tool)."""
import os, subprocess, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from nngp_b200 import _lib
    n, k, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    rng = np.random.default_rng(0)
    a = rng.standard_normal((n, n + 8)); spd = a @ a.T + n * np.eye(n)
    h = _lib.Handle()
    arr = np.array(spd, order="C")
    import ctypes as C
    rc = h._lib.nngp_diag_potrf(h._h, C.c_void_p(arr.ctypes.data), n)
    np.save(out, arr)
    sys.exit(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
W = 256
ks = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else range(1, n // W + 1)
for k in ks:
    outs = {}
    for la in ("0", "1"):
        env = dict(os.environ, NNGP_CHOL_LOOKAHEAD=la, NNGP_POTRF_STOP_AFTER=str(k), NNGP_LA_SAMEPRIO="1", NNGP_CHOL_W=str(W))
        bad = None
        for rep in range(3 if la == "1" else 1):
            f = f"/tmp/potrf_{la}_{rep}.npy"
            subprocess.run([sys.executable, __file__, "child", str(n), str(k), f], env=env, check=True)
            outs[(la, rep)] = np.tril(np.load(f))
    ref = outs[("0", 0)]
    for rep in range(3):
        d = outs[("1", rep)] != ref
        if d.any():
            rows, cols = np.nonzero(d)
            blocks = sorted({(int(r) // 128, int(c) // 64) for r, c in zip(rows, cols)})
            print(f"k={k} rep={rep}: {d.sum()} entries differ; cols {cols.min()}..{cols.max()} rows {rows.min()}..{rows.max()}; first 12 (rowtile128, coltile64): {blocks[:12]}")
            c0 = cols.min(); rr = np.unique(rows[cols == c0])
            print(f"      first col {c0}: rows {rr[:20].tolist()} ... count {len(rr)}")
        else:
            print(f"k={k} rep={rep}: identical")

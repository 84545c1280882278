"""BASELINE config C4: the active-learning loop (active/ActiveLearner.py:67-77) on the B200 path.

Refit-from-scratch as the training set grows n0 -> n_max in steps of `budget`, each round predicting the whole
remaining pool (mean + variance) and selecting the top-`budget` rows by std/max(mean) (the deterministic branch,
ActiveLearner.py:54).  Prints one JSON line with per-round timings.  (That the selected rows are the oracle's is
checked in tests/test_gpu_parity.py.)

    python tools/bench_active.py [--n0 2048 --budget 2048 --n-max 32768 --pool 65536 --dim 128 --depth 2]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "nngp-src_b200"):
    sys.path.insert(0, str(p))
from nngp_b200 import stax, synth  # noqa: E402
from nngp_b200.active import ActiveLearner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n0", type=int, default=2048)
    ap.add_argument("--budget", type=int, default=2048)
    ap.add_argument("--n-max", type=int, default=32768)
    ap.add_argument("--pool", type=int, default=65536)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--depth", type=int, default=2)
    a = ap.parse_args()
    total = a.pool + a.n0
    x = synth.encodings(total, a.dim, 1)
    y = synth.labels(x)[:, None] + 0.05 * np.sin(x[:, :1] / 37.0)      # a label with some structure at D=128
    xtr, ytr, xpool, ypool = x[:a.n0], y[:a.n0], x[a.n0:], y[a.n0:]
    layers = [stax.Dense(512)] + [l for _ in range(a.depth - 1) for l in (stax.Relu(), stax.Dense(1))]
    _, _, kernel_fn = stax.serial(*layers)
    al = ActiveLearner(budget=a.budget, active_iters=0, verbose=False)
    rounds = []
    t_all = time.perf_counter()
    pf, xd, yd = None, None, None
    while True:
        t0 = time.perf_counter()
        # first round: fit from host arrays; later rounds: nngp_append_fit of the selected rows (device-side merge)
        pf = al.train(kernel_fn, xtr, ytr, _reserve=(a.n_max, a.pool)) if pf is None else al.retrain(kernel_fn, pf, xtr, ytr, xd, yd)
        pf.engine()                                          # (first round: the lazy fit happens here)
        t_fit = time.perf_counter() - t0
        idx = al.active_test(pf, xpool)                      # nngp_active_select: predict the pool + top-k on the GPU
        dt = time.perf_counter() - t0
        st = pf.engine().stats()
        pf.engine().stats_reset()                            # the engine is reused across rounds (append + refit)
        rounds.append({"n_train": int(xtr.shape[0]), "pool": int(xpool.shape[0]), "seconds": dt, "fit_call_s": t_fit,
                       "select_call_s": dt - t_fit,
                       "fit_ms": st["fit_total_ms"], "predict_ms": st["pred_total_ms"]})
        if xtr.shape[0] + a.budget > a.n_max or xpool.shape[0] <= a.budget:
            break
        xd, yd = xpool[idx], ypool[idx]
        xtr, ytr, xpool, ypool = al.merge_data(idx, xtr, ytr, xpool, ypool)
    out = {"workload": "C4 active-learning loop", "n0": a.n0, "budget": a.budget, "n_max": a.n_max, "pool": a.pool,
           "dim": a.dim, "depth": a.depth, "rounds": len(rounds), "total_seconds": time.perf_counter() - t_all,
           "sum_fit_ms": sum(r["fit_ms"] for r in rounds), "sum_predict_ms": sum(r["predict_ms"] for r in rounds),
           "per_round": rounds}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

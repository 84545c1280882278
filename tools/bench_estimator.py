"""End-to-end throughput of ``Estimator.predict(query_lines)`` (neuroestimator/estimator/estimator.py:42-61): text
lines in, (mean, std) out -- C++ batch encoder + CUDA posterior, with and without the encode/predict pipeline.
Lines are the committed multi-table fixture lines replicated; the model is N training rows of the same encoding.

    python tools/bench_estimator.py [--lines 524288 --n-train 8192]
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
from nngp_b200.encoder import BatchEncoder  # noqa: E402
from nngp_b200.estimator import Estimator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=524288)
    ap.add_argument("--n-train", type=int, default=8192)
    a = ap.parse_args()
    g = np.load(ROOT / "tests" / "golden" / "encoder_golden.npz", allow_pickle=True)
    enc = BatchEncoder(str(g["schema"]))
    base = [str(l) for l in g["lines"]]
    lines = (base * (a.lines // len(base) + 1))[:a.lines]
    rng = np.random.default_rng(0)
    x_train = enc.encode((base * (a.n_train // len(base) + 1))[:a.n_train])
    x_train = x_train + rng.uniform(0, 1e-3, x_train.shape)          # replicated lines: keep K + lambda I well posed
    y_train = rng.uniform(0, 20, (a.n_train, 1))
    est = Estimator("s", "d", "q", X_train=x_train, Y_train=y_train, nngp_encoder=enc, verbose=False)
    est.load_model()
    out = {"lines": a.lines, "n_train": a.n_train, "dim": int(enc.dim)}
    t0 = time.perf_counter(); x = enc.encode(lines); out["encode_only_s"] = time.perf_counter() - t0
    t0 = time.perf_counter(); est.predict_fn(x_test=x, get="nngp", compute_cov=True); out["predict_only_s"] = time.perf_counter() - t0
    for name, chunk in (("one_shot", 1 << 40), ("pipelined", Estimator.pipeline_chunk)):
        est.pipeline_chunk = chunk
        est.predict(lines[:70000])
        t0 = time.perf_counter()
        mean, std = est.predict(lines)
        dt = time.perf_counter() - t0
        out[name] = {"seconds": dt, "lines_per_s": a.lines / dt}
        out[name + "_checksum"] = float(np.sum(mean) + np.sum(std))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

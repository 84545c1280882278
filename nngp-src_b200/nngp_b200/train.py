"""First-party CLI mirroring the reference's ``train.py`` for the NNGP / NTK path (train.py:225-298):

    python -m nngp_b200.train --kernel_type nngp --relations forest --names forest \\
        --query_path ./Queries/forest_data [--data_path <dir with forest.csv>] [--gpus 8]

Same flags, same printed lines (``number of query``, the shapes, ``Kernel construction in``, ``Mean Square Error``,
``Inference time=``), same data flow -- sorted query files (QuerySampler.py:172-186), ``col,upper,lower#...@card``
lines (:157-170), ``[upper, lower]`` per numerical column scaled to [0, 1000] with the inverted no-predicate default
(:200-221), ``log2(card)`` labels (:188-198), ``random.seed(10)`` 60/20/20 split (util.py:271-293) -- and then
``NNGP_train_and_test`` (train.py:153-203) on the B200 engine.  What differs:
  * the per-line Python encoder loop is the C++ batch encoder (``nngp_encode_lines``, bit-identical rows);
  * the table's per-column (min, max) come from ``<data_path>/<relation>.csv`` when it exists (QuerySampler.py:50-51);
    ``forest.csv`` is not shipped with the reference (readme.md:37), so for ``forest`` the UCI Covertype ranges of
    columns A..J are built in; ``--col_ranges`` (JSON ``{"A": [min, max], ...}``) supplies them for anything else;
  * the report is the symmetric q-error summary instead of ``util.PredictionStatistics`` (seaborn / matplotlib);
  * join queries (``--relations a,b``) need the reference's schema loaders (pandas, out of the hot-path scope): run the
    reference's own ``train.py`` on the import shims instead;
  * ``--gpus G`` predicts on G GPUs of this process (``runtime.set_gpus``), ``--latency`` builds the explicit inverse,
    ``--variance_slices S`` moves the variance product of large batches to the INT8 tensor cores.
The reference's ``--kernel_type gp`` branch (train.py:60-150) is broken upstream (``jit`` undefined) and not mirrored.
"""
from __future__ import annotations

import datetime
import json
import os
import random
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser

import numpy as np

from . import batch as _batch
from . import predict as _predict
from . import runtime, stax
from .encoder import FORMAT_SINGLE_TABLE, BatchEncoder

FOREST_RANGES = {"A": (1859, 3858), "B": (0, 360), "C": (0, 66), "D": (0, 1397), "E": (-173, 601),
                 "F": (0, 7117), "G": (0, 254), "H": (0, 254), "I": (0, 254), "J": (0, 7173)}


def column_ranges(args) -> dict:
    if getattr(args, "col_ranges", None):
        return {k: (float(v[0]), float(v[1])) for k, v in json.loads(args.col_ranges).items()}
    rel = args.relations.split(",")[0]
    csv = os.path.join(args.data_path or "", rel + ".csv")
    if args.data_path and os.path.exists(csv):
        import pandas as pd
        df = pd.read_csv(csv)                                         # QuerySampler.py:30-51
        return {c: (float(df[c].min()), float(df[c].max())) for c in df.columns}
    if rel == "forest":
        return {k: (float(a), float(b)) for k, (a, b) in FOREST_RANGES.items()}
    raise SystemExit(f"no column ranges for relation {rel!r}: pass --data_path with {rel}.csv or --col_ranges")


def load_training_data(args):
    """``datasets.load_training_data`` for a single-table query directory -> (X [n, 2*cols] f64, Y [n, 1] f64)."""
    ranges = column_ranges(args)
    schema = "chunk_size %d\ntable t\n" % min(int(args.chunk_size), 64) + "\n".join(
        f"col {c} num {lo!r} {hi - lo!r}" for c, (lo, hi) in ranges.items())
    enc = BatchEncoder(schema)
    lines = []
    for name in sorted(os.listdir(args.query_path)):                  # QuerySampler.py:176: sorted => query_10 first
        with open(os.path.join(args.query_path, name)) as fh:
            lines += [l for l in fh if l.strip()]
    x, card = enc.encode(lines, fmt=FORMAT_SINGLE_TABLE, with_card=True)
    return x, np.log2(card)[:, None]


def train_test_val_split(X, Y, train_frac=0.6, test_frac=0.2, seed=10):   # util.py:271-293
    n = X.shape[0]
    print("# instances = {}".format(n))
    num_train, num_test = int(train_frac * n), int(test_frac * n)
    indices = list(range(n))
    random.seed(seed)
    random.shuffle(indices)
    X, Y = X[indices, :], Y[indices, :]
    return (X[:num_train], Y[:num_train], X[num_train:num_train + num_test], Y[num_train:num_train + num_test],
            X[num_train + num_test:], Y[num_train + num_test:])


def NNGP_train_and_test(args, X_train, Y_train, X_test, Y_test):         # train.py:153-203
    init_fn, apply_fn, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    kernel_fn = _batch.batch(kernel_fn, device_count=0, batch_size=0)
    start = datetime.datetime.now()
    predict_fn = _predict.gradient_descent_mse_ensemble(kernel_fn, X_train, Y_train, diag_reg=1e-3)
    print('Kernel construction in %s seconds.' % (datetime.datetime.now() - start).total_seconds())
    pred_mean, pred_cov = predict_fn(x_test=X_test, get=args.kernel_type, compute_cov=True)
    pred_std = np.sqrt(np.diag(pred_cov))
    mse = np.sum(np.power(pred_mean - Y_test, 2))
    print("Mean Square Error: {}".format(mse))
    print(X_test.shape, Y_test.shape)
    start = datetime.datetime.now()
    pred_mean, pred_cov = predict_fn(x_test=X_test, get=args.kernel_type, compute_cov=True)
    print("Inference time={} seconds".format((datetime.datetime.now() - start).total_seconds()))
    errors = np.ravel(pred_mean - Y_test)
    q = 2.0 ** np.abs(errors)                                              # util.py:152-167 reports 2**err
    print("q-error: median {:.4f}  mean {:.4f}  95th {:.4f}  max {:.4f}".format(
        float(np.median(q)), float(np.mean(q)), float(np.quantile(q, 0.95)), float(np.max(q))))
    print("posterior std: min {:.4f}  mean {:.4f}  max {:.4f}".format(
        float(np.min(pred_std)), float(np.mean(pred_std)), float(np.max(pred_std))))
    return pred_mean, pred_std, float(mse)


def main(args):
    if args.kernel_type not in ("nngp", "ntk"):
        raise NotImplementedError(f"--kernel_type {args.kernel_type}: 'nngp' and 'ntk' are implemented")
    if args.join_query:
        raise NotImplementedError("join queries go through the reference's schema loaders: run the reference's train.py "
                                  "on the import shims (PYTHONPATH=nngp-src_b200/compat:nngp-src_b200, INTEGRATION.md)")
    if args.gpus > 1:
        runtime.set_gpus(args.gpus)
    if args.latency:
        runtime.set_latency_mode(True)
    if args.variance_slices:
        runtime.set_variance_slices(args.variance_slices)
    X, Y = load_training_data(args)
    print("number of query: {}".format(X.shape[0]))
    X_train, Y_train, X_test, Y_test, _xv, _yv = train_test_val_split(X, Y, train_frac=0.6, test_frac=0.2)
    print(X_train.shape, X_test.shape)
    print(Y_train.shape, Y_test.shape)
    return NNGP_train_and_test(args, X_train, Y_train, X_test, Y_test)


def build_parser() -> ArgumentParser:
    p = ArgumentParser("NNGP/NTK estimator (B200)", formatter_class=ArgumentDefaultsHelpFormatter, conflict_handler="resolve")
    p.add_argument("--chunk_size", default=64, type=int, help="dimension of factorized encoding")
    p.add_argument("--kernel_type", type=str, default="nngp", help="nngp, ntk")
    p.add_argument("--feat_encode", type=str, default="dnn-encoder", help="accepted for compatibility (unused on this path)")
    p.add_argument("--no-cuda", action="store_true", default=True,
                   help="accepted for compatibility: the reference always runs on CPU, this path always on the B200")
    p.add_argument("--relations", type=str, default="forest")
    p.add_argument("--names", type=str, default="forest")
    p.add_argument("--query_path", type=str, default="./Queries/forest_data")
    p.add_argument("--data_path", type=str, default="")
    p.add_argument("--schema_name", type=str, default="imdb_simple", help="accepted for compatibility (join workloads)")
    p.add_argument("--col_ranges", type=str, default="", help='JSON {"col": [min, max], ...} instead of <data_path>/<relation>.csv')
    p.add_argument("--gpus", type=int, default=1, help="GPUs of this process to shard prediction over")
    p.add_argument("--latency", action="store_true", help="latency mode: explicit inverse factor for small batches")
    p.add_argument("--variance_slices", type=int, default=0,
                   help="5..9: variance of large batches on the INT8 tensor cores (digit planes, tcgen05); 0 = all FP64")
    return p


if __name__ == "__main__":
    a = build_parser().parse_args()
    a.join_query = len(a.relations.split(",")) > 1
    print(a)
    main(a)

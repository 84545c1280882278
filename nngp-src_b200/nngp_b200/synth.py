"""Seeded synthetic query encodings shaped like the reference encoders' output.

Single-table (``QuerySampler.py:200-221``): D/2 numeric columns, interleaved ``[upper, lower]`` per
column in [0, 1000]; a column without a predicate carries the encoder's inverted default ``(0, 1000)``.
Multi-join (``JoinQuerySampler.py:604-621``): trailing join-operator one-hot triples.
Seeds: train 1, test 2 (numpy PCG64).  Labels are a deterministic independence-selectivity surrogate
in the log2-cardinality domain -- any deterministic y works, alpha is linear in y.
"""
from __future__ import annotations

import numpy as np


def encodings(rows: int, dim: int, seed: int, pred_prob: float = 0.4, join_dims: int = 0) -> np.ndarray:
    if dim % 2 or join_dims % 3 or join_dims > dim:
        raise ValueError("dim must be even and join_dims a multiple of 3 no larger than dim")
    rng = np.random.default_rng(seed)
    ncol = (dim - join_dims) // 2
    ncol_dims = 2 * ncol
    pad = dim - join_dims - ncol_dims
    a = rng.uniform(0.0, 1000.0, size=(rows, ncol))
    b = rng.uniform(0.0, 1000.0, size=(rows, ncol))
    has = rng.random((rows, ncol)) < pred_prob
    upper = np.where(has, np.maximum(a, b), 0.0)
    lower = np.where(has, np.minimum(a, b), 1000.0)
    x = np.empty((rows, dim), dtype=np.float64)
    x[:, 0:ncol_dims:2] = upper
    x[:, 1:ncol_dims:2] = lower
    if pad:
        x[:, ncol_dims:ncol_dims + pad] = 0.0
    if join_dims:
        trip = join_dims // 3
        on = rng.random((rows, trip)) < 0.15
        j = np.zeros((rows, trip, 3))
        j[:, :, 0] = on  # one-hot on the '=' slot
        x[:, dim - join_dims:] = j.reshape(rows, join_dims)
    return x


def labels(x: np.ndarray, join_dims: int = 0) -> np.ndarray:
    ncol_dims = (x.shape[1] - join_dims) // 2 * 2
    up, lo = x[:, 0:ncol_dims:2], x[:, 1:ncol_dims:2]
    has = ~((up == 0.0) & (lo == 1000.0))
    sel = np.where(has, np.maximum(up - lo, 1.0) / 1000.0, 1.0)
    y = np.log2(1.0 + 5.8e5 * np.prod(sel, axis=1))
    return np.clip(y, 0.0, 20.0)


def make_problem(n_train: int, n_test: int, dim: int, join_dims: int = 0, seed_train: int = 1, seed_test: int = 2):
    xtr = encodings(n_train, dim, seed_train, join_dims=join_dims)
    xte = encodings(n_test, dim, seed_test, join_dims=join_dims)
    return xtr, labels(xtr, join_dims), xte, labels(xte, join_dims)

"""``jax.random`` surface of active/ActiveLearner.py:50-52 (biased sampling).  JAX's threefry stream is not
reproducible without jax, so this is numpy's PCG64 Gumbel-free weighted choice: same distribution, different
draws (that branch is unpinned; parity uses the deterministic argsort branch)."""
import numpy as _np


def PRNGKey(seed):
    return int(seed)


def choice(key, a, shape=(), replace=True, p=None):
    rng = _np.random.default_rng(key)
    n = int(a) if _np.ndim(a) == 0 else len(a)
    idx = rng.choice(n, size=shape, replace=replace, p=None if p is None else _np.asarray(p, dtype=_np.float64))
    return idx if _np.ndim(a) == 0 else _np.asarray(a)[idx]

"""Minimal ``jax`` import shim for the reference's NNGP drivers (train.py:7-16, estimator.py:1-13,
active/active_train.py:6-9).  Arrays are plain numpy float64 on the host; all NNGP arithmetic happens in
libnngp_b200.so behind the neural_tangents shim.  Nothing here computes a kernel or a solve."""
from . import numpy, random, scipy  # noqa: F401
from .config import config  # noqa: F401


def _not_on_hot_path(name):
    def fn(*_a, **_k):
        raise NotImplementedError(f"jax.{name} is not used on the NNGP hot path and is not provided by this shim")
    fn.__name__ = name
    return fn


def jit(fn=None, **_k):
    return fn if fn is not None else (lambda f: f)


grad = _not_on_hot_path("grad")
vmap = _not_on_hot_path("vmap")
pmap = _not_on_hot_path("pmap")

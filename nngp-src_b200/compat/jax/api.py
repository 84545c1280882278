from . import jit  # noqa: F401   (active/ActiveLearner.py:7 imports jax.api.jit and never calls it)

"""``import nngp_b200.nt as nt`` -- the ``neural_tangents`` top-level names the reference uses:
``nt.batch`` (train.py:166) and ``nt.predict.gradient_descent_mse_ensemble`` (train.py:171)."""
from . import predict, stax  # noqa: F401
from .batch import batch  # noqa: F401

"""Import shim: ``import neural_tangents as nt`` / ``from neural_tangents import stax`` resolve to the
B200 path, so the reference's drivers (train.py:16-17,161-172; neuroestimator/estimator/estimator.py:7-8;
active/ActiveLearner.py:9-10) run unmodified with ``PYTHONPATH=nngp-src_b200/compat:nngp-src_b200``.
Only the surface those call sites touch exists; everything else raises NotImplementedError."""
from nngp_b200 import predict, stax  # noqa: F401
from nngp_b200.batch import batch  # noqa: F401

__version__ = "0.6.1+nngp_b200"

"""``jax.numpy`` -> numpy passthrough.  ``np.diag(pred_cov)`` (train.py:180, estimator.py:55,
ActiveLearner.py:46) reaches ``LazyCovariance.__array_function__`` and returns the posterior-variance
diagonal without materialising T x T."""
from numpy import *  # noqa: F401,F403
from numpy import asarray, diag, float64, sqrt  # noqa: F401

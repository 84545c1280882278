"""Mirror of ``active/ActiveLearner.py:14-77`` (the NNGP active-learning loop, BASELINE config C4).

Refit-from-scratch on a growing training set with posterior-variance query selection (SURVEY.md section 8a row
a8, 8f-3).  The loop stays Python as in the reference; the two steps inside it run on the device through the C ABI:
  * ``active_test``  -> ``nngp_active_select``: predict the pool, score = std / max(mean), exact top-k (or
    Gumbel-top-k sampling) by radix select on the GPU; only the selected row numbers come back;
  * ``train`` after a merge -> ``nngp_append_fit``: the selected rows are appended to the training set that already
    sits in HBM and the model is refit (same arithmetic as a fresh fit -- bitwise -- without re-sending X_train).
Differences, both stated in DESIGN.md:
  * the biased-sampling branch draws from a splitmix64 stream instead of JAX's threefry stream
    (``jax.random.choice`` is not reproducible without jax); the deterministic top-k branch is the one parity is
    checked on;
  * reporting is the symmetric q-error summary instead of ``util.PredictionStatistics`` (plotting deps).
"""
from __future__ import annotations

import numpy as np

from . import batch as _batch
from . import predict as _predict


class ActiveLearner(object):
    def __init__(self, args=None, budget=1000, active_iters=3, kernel_type="nngp", biased_sample=False, verbose=True):
        self.args = args
        self.budget = getattr(args, "budget", budget)
        self.active_iters = getattr(args, "active_iters", active_iters)
        self.kernel_type = getattr(args, "kernel_type", kernel_type)
        self.biased_sample = getattr(args, "biased_sample", biased_sample)
        self._say = print if verbose else (lambda *a, **k: None)
        self.history = []

    def train(self, kernel_fn, X_train, Y_train, X_test=None, Y_test=None, _reserve=None):  # ActiveLearner.py:23-31
        kernel_fn = _batch.batch(kernel_fn, device_count=0, batch_size=0)
        predict_fn = _predict.gradient_descent_mse_ensemble(kernel_fn, X_train, Y_train, diag_reg=1e-3,
                                                            _reserve=_reserve)
        if X_test is not None and Y_test is not None:
            self.test(predict_fn, X_test, Y_test, None)
        return predict_fn

    def retrain(self, kernel_fn, predict_fn, X_train, Y_train, X_delta, Y_delta):
        """``self.train(kernel_fn, X_train, Y_train)`` (ActiveLearner.py:76) where X_train = [old; X_delta]: the fitted
        engine of ``predict_fn`` appends the new rows on the device and refits; the returned predict_fn is bound to it.
        The engine is TAKEN from ``predict_fn``: a caller that keeps the old closure (to compare rounds) gets the old
        model back -- it refits lazily from its own data -- never the new one."""
        if not hasattr(predict_fn, "engine"):
            return self.train(kernel_fn, X_train, Y_train)
        h = predict_fn.engine(self.kernel_type)
        if hasattr(predict_fn, "release"):     # the engine changes owner: the old closure refits if it is used again
            predict_fn.release(self.kernel_type)
        h.append_fit(X_delta, Y_delta)
        kernel_fn = _batch.batch(kernel_fn, device_count=0, batch_size=0)
        return _predict.gradient_descent_mse_ensemble(kernel_fn, X_train, Y_train, diag_reg=1e-3,
                                                      _fitted_engines={self.kernel_type: h})

    def test(self, predict_fn, X_val, Y_val, query_infos_val=None, kernel_type="nngp", compute_cov=True):  # :33-40
        pred_mean, pred_cov = predict_fn(x_test=X_val, get=kernel_type, compute_cov=compute_cov)
        errors = pred_mean - Y_val
        mse = float(np.mean(np.power(errors, 2.0)))
        q = 2.0 ** np.abs(np.asarray(errors).ravel())
        rec = {"mse": mse, "qerr_median": float(np.median(q)), "qerr_p95": float(np.quantile(q, 0.95)),
               "qerr_max": float(np.max(q))}
        self.history.append(rec)
        self._say("Test MSE Loss:{}".format(mse))
        return rec

    def active_test(self, predict_fn, X_test, kernel_type="nngp"):                          # ActiveLearner.py:43-55
        if hasattr(predict_fn, "engine"):      # device path: selection happens next to the posterior, on the GPU
            return predict_fn.engine(kernel_type).active_select(X_test, self.budget, self.biased_sample, seed=10)
        return self._active_test_host(predict_fn, X_test, kernel_type)

    def _active_test_host(self, predict_fn, X_test, kernel_type="nngp"):
        """The same selection with numpy on the host (any predict_fn; used by the tests as a cross-check)."""
        pred_mean, pred_cov = predict_fn(x_test=X_test, get=kernel_type, compute_cov=True)
        pred_std = np.sqrt(np.diag(pred_cov))
        pred_std = pred_std / np.max(pred_mean, 0)
        num_test = X_test.shape[0]
        pred_std = np.reshape(pred_std, (num_test,))
        num_select = self.budget if num_test > self.budget else num_test
        if self.biased_sample:
            std_prob = pred_std / np.sum(pred_std)
            return np.random.default_rng(10).choice(num_test, size=(num_select,), replace=False, p=std_prob)
        return np.argsort(pred_std)[-num_select:]

    def merge_data(self, select_indices, X_train, Y_train, X_test, Y_test):                 # ActiveLearner.py:57-65
        X_delta, Y_delta = X_test[select_indices], Y_test[select_indices]
        X_train_new = np.vstack((X_train, X_delta))
        Y_train_new = np.vstack((Y_train, Y_delta))
        keep = np.setdiff1d(np.arange(X_test.shape[0]), np.asarray(select_indices))
        return X_train_new, Y_train_new, X_test[keep], Y_test[keep]

    def active_train(self, kernel_fn, X_train, Y_train, X_test, Y_test, X_val, Y_val, query_infos_val=None):  # :67-77
        self._say("# Initial Training samples: {}".format(X_train.shape[0]))
        # the loop's final sizes are known up front: size the device buffers once (nngp_reserve)
        grow = min(self.active_iters * self.budget, X_test.shape[0])
        predict_fn = self.train(kernel_fn, X_train, Y_train,
                                _reserve=(X_train.shape[0] + grow, max(X_test.shape[0], X_val.shape[0])))
        self.test(predict_fn, X_val, Y_val, query_infos_val, self.kernel_type)
        for i in range(self.active_iters):
            select_indices = self.active_test(predict_fn, X_test, self.kernel_type)
            self._say("Active Iteration {}: Selection {}".format(i, select_indices.shape[0]))
            X_delta, Y_delta = X_test[select_indices], Y_test[select_indices]
            X_train, Y_train, X_test, Y_test = self.merge_data(select_indices, X_train, Y_train, X_test, Y_test)
            self._say("# Training samples: {}".format(X_train.shape[0]))
            predict_fn = self.retrain(kernel_fn, predict_fn, X_train, Y_train, X_delta, Y_delta)
            self.test(predict_fn, X_val, Y_val, query_infos_val, self.kernel_type)
        return predict_fn, X_train, Y_train

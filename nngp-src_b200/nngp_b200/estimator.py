"""``neuroestimator.Estimator`` with the NNGP arithmetic on the B200 path.

Mirror of neuroestimator/estimator/estimator.py:16-68 (class signature, ``load_model``, ``predict`` return
types).  Loading the schema / training queries and encoding query lines are host string/pandas work that is
out of the hot-path scope, so they are injected: pass either the
reference's own ``load_training_schema_data`` (its signature is kept) or pre-encoded arrays + an encoder.
"""
from __future__ import annotations

import datetime
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import batch as _batch
from . import predict as _predict
from . import stax


class Estimator(object):
    # query lines per pipeline stage of predict(): one stage is one full-efficiency GPU batch (the persistent solve
    # wants >= 2 x 148 row tiles of 128); the next stage's lines are parsed on the host cores meanwhile
    pipeline_chunk = 65536

    def __init__(self, schema_name: str = "", data_path: str = "", train_query_path: str = "", chunk_size: int = 64,
                 use_aux: bool = False, q_error_threshold: float = 100.0, coef_var_threshold: float = 1.0, *,
                 loader=None, X_train=None, Y_train=None, nngp_encoder=None, verbose: bool = True):
        self.schema_name, self.data_path = schema_name, data_path
        self.train_query_path, self.chunk_size = train_query_path, chunk_size
        self._say = print if verbose else (lambda *a, **k: None)
        if X_train is None:
            if loader is None:
                raise ValueError("Estimator: pass loader=<the reference's load_training_schema_data> "
                                 "(neuroestimator/estimator/util.py:159-195) or X_train/Y_train/nngp_encoder")
            self._say("loading schema and training data ... This may take seconds ...")
            X_train, Y_train, nngp_encoder = loader(schema_name, data_path, train_query_path, chunk_size, use_aux,
                                                    q_error_threshold, coef_var_threshold)
        self.nngp_encoder = nngp_encoder
        self.X_train, self.Y_train = np.asarray(X_train, dtype=np.float64), np.asarray(Y_train, dtype=np.float64)
        self._say("Building model kernel ...")
        init_fn, apply_fn, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))   # estimator.py:27-30
        kernel_fn = _batch.batch(kernel_fn, device_count=0, batch_size=0)                         # estimator.py:31-33
        self.predict_fn = _predict.gradient_descent_mse_ensemble(kernel_fn, self.X_train, self.Y_train,
                                                                 diag_reg=1e-3)                  # estimator.py:34-35

    def load_model(self):
        """Forces the fit (the reference's warm-up predicts the training set, estimator.py:37-40)."""
        pred_mean, pred_cov = self._nngp_prediction(self.X_train)
        self._say(pred_mean.shape, pred_cov.shape)
        self._say("Model construction complete.")

    def predict(self, query_lines):
        start = datetime.datetime.now()
        if self.nngp_encoder is None:
            raise ValueError("Estimator.predict needs nngp_encoder.parse_line_without_card_then_encode")
        if hasattr(self.nngp_encoder, "encode"):        # nngp_b200.encoder.BatchEncoder: all lines in one C++ call
            if len(query_lines) > self.pipeline_chunk:  # ... or chunk i+1 parsed while the GPU predicts chunk i
                return self._predict_pipelined(list(query_lines), start)
            X_test = self.nngp_encoder.encode(query_lines)
        else:                                           # the reference's per-line loop, estimator.py:46-50
            X_test = np.asarray([self.nngp_encoder.parse_line_without_card_then_encode(line) for line in query_lines],
                                dtype=np.float64)
        pred_mean, pred_cov = self._nngp_prediction(X_test)
        duration = (datetime.datetime.now() - start).total_seconds()
        self._say("prediction time={} seconds".format(duration))
        pred_std = np.sqrt(np.diag(pred_cov))                                                      # estimator.py:55
        return pred_mean.ravel(), pred_std.ravel()                                                 # estimator.py:56,61

    def _predict_pipelined(self, query_lines, start):
        """Large batches: the C++ line encoder (host cores, GIL released) runs one chunk ahead of the GPU.  Rows are
        independent and the CUDA path is bitwise invariant to how rows are blocked, so the result equals the
        one-shot path's; the two stages cost max(encode, predict) per chunk instead of their sum."""
        step = self.pipeline_chunk
        chunks = [query_lines[i:i + step] for i in range(0, len(query_lines), step)]
        means, stds = [], []
        with ThreadPoolExecutor(max_workers=1) as pool:
            pending = pool.submit(self.nngp_encoder.encode, chunks[0])
            for i in range(len(chunks)):
                X_test = pending.result()
                if i + 1 < len(chunks):
                    pending = pool.submit(self.nngp_encoder.encode, chunks[i + 1])
                pred_mean, pred_cov = self._nngp_prediction(X_test)
                means.append(np.asarray(pred_mean).ravel())
                stds.append(np.sqrt(np.diag(pred_cov)).ravel())
        duration = (datetime.datetime.now() - start).total_seconds()
        self._say("prediction time={} seconds".format(duration))
        return np.concatenate(means), np.concatenate(stds)

    def _nngp_prediction(self, X_test, kernel_type="nngp", compute_cov=True):
        return self.predict_fn(x_test=X_test, get=kernel_type, compute_cov=compute_cov)           # estimator.py:64-68

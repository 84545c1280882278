"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED AGAINST THE REFERENCE.

A numpy/scipy FP64 restatement of the one hot path of Kangfei/NNGP-src that this repo replaces:
NNGP kernel construction + exact GP posterior inference.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / CPU baseline -- never the product path
(``nngp_b200`` never imports it and has no CPU fallback).

Why "parity unpinned": the reference delegates ALL arithmetic on this path to the third-party
package **neural-tangents 0.6.1** on **jax 0.3.23 / jaxlib 0.3.22** (pins: reference
``nngp.yaml:78,79,88``), which is neither vendored under the reference tree nor installable here
(no network), and the reference ships no tests, golden vectors or fixtures for the path
(SURVEY.md section 4 / 8c).  This file therefore restates neural-tangents' published algorithm and is
anchored on the reference's own call sites; it is pinned instead by (i) analytic known answers,
(ii) a 50-digit mpmath restatement (``oracle/mp_oracle.py`` -> ``tests/golden/mp_small.npz``) and
(iii) a finite-width Monte-Carlo check of the conventions (``tests/test_oracle.py``).

Reference call sites restated (file:line under the reference tree):
  * model        stax.serial(Dense(512), Relu(), Dense(1))            train.py:161-164,
                 neuroestimator/estimator/estimator.py:27-30, active/active_train.py:40-43
  * FP64         config.update("jax_enable_x64", True)                train.py:24, estimator.py:12
  * fit          nt.predict.gradient_descent_mse_ensemble(kernel_fn, X, Y, diag_reg=1e-3)
                                                                      train.py:171-172, estimator.py:34-35,
                                                                      active/ActiveLearner.py:27-28
  * predict      predict_fn(x_test=X, get='nngp', compute_cov=True)   train.py:157-158, estimator.py:66-67
  * std          np.sqrt(np.diag(pred_cov))                           train.py:180, estimator.py:55,
                                                                      active/ActiveLearner.py:46
  * AL selection std/max(mean); argsort(std)[-budget:]                active/ActiveLearner.py:46-54
[nt-upstream] callees restated: stax._inputs_to_kernel, stax.Dense (_affine), stax.ABRelu(a=0,b=1)
(0.6.1 arctan2 form), predict._add_diagonal_regularizer, predict._get_cho_solve, predict.gp_inference.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.linalg as sla

TWO_PI = 2.0 * np.pi
# threads for the elementwise arc-cosine recursion (the BLAS calls use OpenBLAS' own pool)
THREADS = int(os.environ.get("NNGP_ORACLE_THREADS", "0")) or len(os.sched_getaffinity(0))
_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=THREADS)
    return _POOL


def _layer_sigmas(depth, sigma_w, sigma_b):
    """(sw2[l], sb2[l]) for Dense layer l = 0..depth-1: scalars apply to every layer, sequences give one value per
    layer [nt: every stax.Dense carries its own W_std / b_std]."""
    sw = np.broadcast_to(np.asarray(sigma_w, dtype=np.float64), (depth,)) if np.ndim(sigma_w) == 0 else np.asarray(sigma_w, dtype=np.float64)
    sb = np.broadcast_to(np.asarray(sigma_b, dtype=np.float64), (depth,)) if np.ndim(sigma_b) == 0 else np.asarray(sigma_b, dtype=np.float64)
    if sw.shape != (depth,) or sb.shape != (depth,):
        raise ValueError(f"per-layer sigma_w / sigma_b need {depth} values")
    return sw ** 2, sb ** 2


def _relu_layers(k, ntk, q1, q2, depth, sw2l, sb2l):
    """depth-1 ReLU arc-cosine steps on the block k (rows: q1, columns: q2), in place (Appendix A.1 / A.5);
    sw2l[l], sb2l[l]: Dense layer l (the step with index l-1 is followed by layer l)."""
    factor = 1.0 / TWO_PI
    q1, q2l = q1.copy(), q2.copy()
    for layer in range(1, depth):
        sw2, sb2 = sw2l[layer], sb2l[layer]
        prod = q1[:, None] * q2l[None, :]
        s = np.sqrt(np.maximum(prod - k * k, 0.0))
        theta = np.arctan2(s, k)
        theta[(s == 0.0) & (k == 0.0)] = np.pi / 2
        dot_sigma = 0.5 - factor * theta
        k[...] = sw2 * (factor * s + dot_sigma * k) + sb2
        if ntk is not None:
            ntk[...] = k + sw2 * (ntk * dot_sigma)
        q1 = sw2 * (0.5 * q1) + sb2
        q2l = sw2 * (0.5 * q2l) + sb2


def layer0_diag(x: np.ndarray, sigma_w=1.0, sigma_b=0.0) -> np.ndarray:
    """q0 = sigma_w^2 |x|^2 / D + sigma_b^2   [nt: _inputs_to_kernel (/channel count) then Dense._affine]
    (per-layer sequences: the first layer's values)."""
    x = np.asarray(x, dtype=np.float64)
    sw = float(np.asarray(sigma_w, dtype=np.float64).reshape(-1)[0])
    sb = float(np.asarray(sigma_b, dtype=np.float64).reshape(-1)[0])
    return sw**2 * (np.einsum("ij,ij->i", x, x) / x.shape[1]) + sb**2


def final_diag(q0: np.ndarray, depth: int = 2, sigma_w=1.0, sigma_b=0.0) -> np.ndarray:
    """K(x,x) after depth-1 ReLU steps: q <- sigma_w^2 q/2 + sigma_b^2   [nt: ABRelu nngp_fn_diag + Dense]."""
    sw2l, sb2l = _layer_sigmas(depth, sigma_w, sigma_b)
    q = np.array(q0, dtype=np.float64, copy=True)
    for layer in range(1, depth):
        q = sw2l[layer] * (0.5 * q) + sb2l[layer]
    return q


def kernel_fn(x1: np.ndarray, x2: np.ndarray | None = None, depth: int = 2, sigma_w: float = 1.0,
              sigma_b: float = 0.0, chunk: int = 2048, get: str = "nngp"):
    """NNGP kernel K(x1, x2) -- or, with get='ntk', the NTK Theta(x1, x2) -- of a depth-`depth` Dense/ReLU
    network.

    k0 = sigma_w^2 x1.x2^T / D + sigma_b^2, theta_ntk0 = k0, then depth-1 times
      s = sqrt(max(q1 q2 - k^2, 0)); theta = arctan2(s, k) (pi/2 where s == k == 0); kdot = 1/2 - theta/(2 pi)
      k <- sigma_w^2 ( s/(2 pi) + kdot k ) + sigma_b^2 ;  q <- sigma_w^2 q/2 + sigma_b^2
      ntk <- k + sigma_w^2 (ntk kdot)                      [nt: Relu `ntk *= dot_sigma`, Dense `ntk = nngp + W_std^2 ntk`]
    get: 'nngp' | 'ntk' | 'both' (returns (K, Theta)).  Row-chunked so the temporaries stay bounded.
    """
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = x1 if x2 is None else np.asarray(x2, dtype=np.float64)
    D = x1.shape[1]
    sw2l, sb2l = _layer_sigmas(depth, sigma_w, sigma_b)
    sw2, sb2 = sw2l[0], sb2l[0]
    q1_all, q2 = layer0_diag(x1, sigma_w, sigma_b), layer0_diag(x2, sigma_w, sigma_b)
    out = np.empty((x1.shape[0], x2.shape[0]), dtype=np.float64)
    out_ntk = np.empty_like(out) if get in ("ntk", "both") else None
    for r0 in range(0, x1.shape[0], chunk):
        r1 = min(r0 + chunk, x1.shape[0])
        k = sw2 * ((x1[r0:r1] @ x2.T) / D) + sb2          # BLAS dgemm (all cores)
        ntk = k.copy() if out_ntk is not None else None
        # The elementwise recursion is one fused, multi-threaded loop in XLA:CPU (the reference's backend); numpy's
        # ufuncs are single-threaded but release the GIL, so row slices of the chunk run on a thread pool.  Every
        # entry sees exactly the same operations as in a single pass: results do not depend on the thread count.
        nthr = max(1, min(THREADS, (r1 - r0 + 63) // 64))
        bounds = [(r1 - r0) * i // nthr for i in range(nthr + 1)]

        def work(i, k=k, ntk=ntk, r0=r0):
            a, b = bounds[i], bounds[i + 1]
            if b > a:
                _relu_layers(k[a:b], None if ntk is None else ntk[a:b], q1_all[r0 + a:r0 + b], q2, depth, sw2l, sb2l)

        if nthr == 1:
            work(0)
        else:
            list(_pool().map(work, range(nthr)))
        out[r0:r1] = k
        if out_ntk is not None:
            out_ntk[r0:r1] = ntk
    if get == "nngp":
        return out
    return out_ntk if get == "ntk" else (out, out_ntk)


def final_diag_ntk(q0: np.ndarray, depth: int = 2, sigma_w=1.0, sigma_b=0.0) -> np.ndarray:
    """Theta(x,x): theta = 0 on the diagonal so kdot = 1/2."""
    sw2l, sb2l = _layer_sigmas(depth, sigma_w, sigma_b)
    q = np.array(q0, dtype=np.float64, copy=True)
    ntk = q.copy()
    for layer in range(1, depth):
        q = sw2l[layer] * (0.5 * q) + sb2l[layer]
        ntk = q + sw2l[layer] * (0.5 * ntk)
    return ntk


class Fit:
    """What the first predict_fn call computes and caches in the reference (Appendix A.2)."""

    def __init__(self, x, y, depth=2, sigma_w=1.0, sigma_b=0.0, diag_reg=1e-3, diag_reg_absolute=False):
        self.x = np.asarray(x, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64).reshape(-1)
        self.depth, self.sigma_w, self.sigma_b = depth, sigma_w, sigma_b
        n = self.x.shape[0]
        k = kernel_fn(self.x, None, depth, sigma_w, sigma_b)
        reg = max(diag_reg, 0.0)
        self.lam = reg if diag_reg_absolute else reg * (np.trace(k) / n)   # [nt: _add_diagonal_regularizer]
        k[np.diag_indices(n)] += self.lam
        # [nt: cho_factor]  k is symmetric, so its transpose view (Fortran order) is the same matrix: LAPACK dpotrf
        # then factors in place instead of scipy first copying the 8*N^2 bytes into column-major order
        self.c = sla.cholesky(k.T, lower=True, overwrite_a=True, check_finite=False)
        self.alpha = sla.cho_solve((self.c, True), self.y, check_finite=False)      # [nt: cho_solve]

    def log_marginal_likelihood(self):
        """-1/2 y^T (K+lam I)^-1 y - sum log C_ii - N/2 log 2 pi  (standard GP evidence; cf. train.py:86-103)."""
        n = self.y.shape[0]
        return float(-0.5 * self.y @ self.alpha - np.sum(np.log(np.diag(self.c))) - 0.5 * n * np.log(2 * np.pi))

    def predict(self, x_test, want_var=True, chunk=4096):
        """mean = K_* alpha ; var_i = K(x_i,x_i) - ||C^-1 K_*[i,:]^T||^2 (Appendix A.3, diagonal only)."""
        x_test = np.asarray(x_test, dtype=np.float64)
        t = x_test.shape[0]
        mean = np.empty(t)
        var = np.empty(t) if want_var else None
        for r0 in range(0, t, chunk):
            r1 = min(r0 + chunk, t)
            ks = kernel_fn(x_test[r0:r1], self.x, self.depth, self.sigma_w, self.sigma_b)
            mean[r0:r1] = ks @ self.alpha
            if want_var:
                v = sla.solve_triangular(self.c, ks.T, lower=True, check_finite=False, overwrite_b=True)
                kss = final_diag(layer0_diag(x_test[r0:r1], self.sigma_w, self.sigma_b), self.depth,
                                 self.sigma_w, self.sigma_b)
                var[r0:r1] = kss - np.einsum("ij,ij->j", v, v)
        return mean, var

    def predict_full_cov(self, x_test):
        """The reference's literal output: (mean (T,1), cov (T,T))  [nt: gp_inference.predict_fn]."""
        x_test = np.asarray(x_test, dtype=np.float64)
        ks = kernel_fn(x_test, self.x, self.depth, self.sigma_w, self.sigma_b)
        ktt = kernel_fn(x_test, None, self.depth, self.sigma_w, self.sigma_b)
        mean = ks @ self.alpha
        cov = ktt - ks @ sla.cho_solve((self.c, True), ks.T, check_finite=False)
        return mean[:, None], cov


def active_select(mean, std, budget: int):
    """Deterministic branch of ActiveLearner.active_test (active/ActiveLearner.py:46-54, biased_sample=False)."""
    s = np.asarray(std).reshape(-1) / np.max(mean, 0)
    num = budget if s.shape[0] > budget else s.shape[0]
    return np.argsort(s)[-num:]


def splitmix_uniform(seed: int, n: int):
    """u_i in (0,1), i = 0..n-1: splitmix64 of (seed, i) -- the counter-based stream the device sampler uses."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.arange(n, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(12)).astype(np.float64) + 0.5) * 2.0 ** -52     # 52 bits: the +0.5 is exact


def active_sample(mean, std, budget: int, seed: int = 10):
    """Biased-sampling branch of ActiveLearner.active_test (active/ActiveLearner.py:49-53, biased_sample=True):
    `budget` draws without replacement with p_i = s_i / sum(s), restated as Gumbel-top-k (the construction
    jax.random.choice itself uses): the rows with the largest log(s_i) - log(-log(u_i)), best first.  The uniform
    stream is splitmix64, not JAX's threefry, so this branch is UNPINNED against the reference."""
    s = np.asarray(std).reshape(-1) / np.max(mean, 0)
    if not np.all(np.isfinite(s)) or np.any(s < 0):
        raise ValueError("sampling needs finite, non-negative scores")
    num = budget if s.shape[0] > budget else s.shape[0]
    u = splitmix_uniform(seed, s.shape[0])
    with np.errstate(divide="ignore"):
        key = np.where(s > 0, np.log(s) - np.log(-np.log(u)), -np.inf)
    order = np.lexsort((np.arange(s.shape[0]), key))      # ascending by (key, index)
    return order[::-1][:num]


def q_error_stats(pred_log2, true_log2):
    """Symmetric q-error 2**|err| summary (util.py:152-167 prints the signed ratio 2**err)."""
    qe = 2.0 ** np.abs(np.asarray(pred_log2).reshape(-1) - np.asarray(true_log2).reshape(-1))
    return {"median": float(np.median(qe)), "mean": float(np.mean(qe)), "p95": float(np.quantile(qe, 0.95)),
            "max": float(np.max(qe))}


class FitNTK:
    """predict_fn(get='ntk') of nt.predict.gradient_descent_mse_ensemble at t = infinity (SURVEY.md Appendix A.5;
    reference flag: train.py:254 `--kernel_type ntk`, consumed at train.py:157-158,178).
      lambda = diag_reg * trace(Theta_dd)/N ; A = Theta_dd + lambda I = C C^T ; alpha = A^-1 y
      mean   = Theta_* alpha
      cov    = K_** + Theta_* A^-1 K_dd A^-1 Theta_*^T - (Theta_* A^-1 K_*^T + transpose)       (diagonal only)
    """

    def __init__(self, x, y, depth=2, sigma_w=1.0, sigma_b=0.0, diag_reg=1e-3, diag_reg_absolute=False):
        self.x = np.asarray(x, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64).reshape(-1)
        self.depth, self.sigma_w, self.sigma_b = depth, sigma_w, sigma_b
        n = self.x.shape[0]
        self.k_dd, th = kernel_fn(self.x, None, depth, sigma_w, sigma_b, get="both")
        reg = max(diag_reg, 0.0)
        self.lam = reg if diag_reg_absolute else reg * (np.trace(th) / n)
        th[np.diag_indices(n)] += self.lam
        self.c = sla.cholesky(th, lower=True, overwrite_a=True, check_finite=False)
        self.alpha = sla.cho_solve((self.c, True), self.y, check_finite=False)

    def predict(self, x_test, want_var=True, chunk=2048):
        x_test = np.asarray(x_test, dtype=np.float64)
        t = x_test.shape[0]
        mean = np.empty(t)
        var = np.empty(t) if want_var else None
        for r0 in range(0, t, chunk):
            r1 = min(r0 + chunk, t)
            ks, ths = kernel_fn(x_test[r0:r1], self.x, self.depth, self.sigma_w, self.sigma_b, get="both")
            mean[r0:r1] = ths @ self.alpha
            if want_var:
                w = sla.cho_solve((self.c, True), ths.T, check_finite=False)          # A^-1 Theta_*^T   (N x rows)
                kss = final_diag(layer0_diag(x_test[r0:r1], self.sigma_w, self.sigma_b), self.depth, self.sigma_w,
                                 self.sigma_b)
                quad = np.einsum("ij,ij->j", w, self.k_dd @ w)
                cross = np.einsum("ij,ji->j", w, ks)
                var[r0:r1] = kss + quad - 2.0 * cross
        return mean, var

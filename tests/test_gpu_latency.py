"""GPU: latency mode (cfg.latency_mode) -- the fit also builds the explicit inverse factor L^-1 and small prediction
batches (the serving case, neuroestimator/estimator/estimator.py:42-62: a few query lines per call) run as a
dependency-free triangular GEMM against it.  Same posterior as the substitution path to rounding (not bitwise), so it
is an opt-in and sits behind the same 1e-6 gate against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import nngp_oracle as oracle  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    from nngp_b200 import _lib
    _lib.load()
    return _lib


@pytest.fixture(scope="module")
def synth():
    from nngp_b200 import synth as s
    return s


def relmax(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.mark.parametrize("n,d,depth", [(130, 8, 2), (1000, 24, 2), (2500, 64, 3), (4097, 33, 2)])
def test_latency_mode_matches_oracle_and_substitution(lib, synth, n, d, depth):
    d2 = d + (d % 2)
    xtr, ytr, xte, _ = synth.make_problem(n, 700, d2)
    xtr, xte = xtr[:, :d], xte[:, :d]
    h = lib.Handle(depth=depth)
    hl = lib.Handle(depth=depth, latency_mode=True)
    h.fit(xtr, ytr)
    hl.fit(xtr, ytr)
    assert hl.stats()["inverse_ms"] > 0
    ref = oracle.Fit(xtr, ytr, depth)
    for rows in (1, 3, 128, 129, 700):
        m, v = h.predict(xte[:rows])
        ml, vl = hl.predict(xte[:rows])
        rm, rv = ref.predict(xte[:rows])
        assert np.array_equal(ml, m)                                   # the mean never touches the factor
        assert np.max(np.abs(vl - rv) / np.abs(rv)) < 1e-6 and np.max(np.abs(v - rv) / np.abs(rv)) < 1e-6
        assert np.max(np.abs(vl - v) / np.abs(v)) < 1e-8
    st = hl.get_state()
    assert relmax(st["l"], ref.c) < 1e-9


def test_latency_mode_threshold_and_refits(lib, synth, monkeypatch):
    """Above NNGP_LATENCY_ROWS the ordinary persistent solve runs (bitwise the default handle's result); appends and
    imports rebuild the inverse."""
    xtr, ytr, xte, _ = synth.make_problem(1500, 6000, 32)
    h = lib.Handle()
    hl = lib.Handle(latency_mode=True)
    h.fit(xtr, ytr)
    hl.fit(xtr, ytr)
    m, v = h.predict(xte)
    ml, vl = hl.predict(xte)                       # 6000 rows > 4096: substitution path
    assert np.array_equal(ml, m) and np.array_equal(vl, v)
    monkeypatch.setenv("NNGP_LATENCY_ROWS", "100000")
    _, vl2 = hl.predict(xte)                       # forced onto the inverse: 47 row tiles
    assert not np.array_equal(vl2, v) and np.max(np.abs(vl2 - v) / np.abs(v)) < 1e-8
    monkeypatch.delenv("NNGP_LATENCY_ROWS")
    xn = synth.encodings(300, 32, 9)
    yn = synth.labels(xn)
    h.append_fit(xn, yn)
    hl.append_fit(xn, yn)
    _, va = h.predict(xte[:50])
    _, vla = hl.predict(xte[:50])
    assert np.max(np.abs(vla - va) / np.abs(va)) < 1e-8 and not np.array_equal(va, v[:50])
    st = h.get_state()
    hi = lib.Handle(latency_mode=True)
    hi.set_state(st["x"], st["l"], st["alpha"], st["lambda"])
    _, vi = hi.predict(xte[:50])
    assert np.array_equal(vi, vla)


def test_latency_mode_on_the_forest_workload(lib, forest):
    """Config C1 (cond(K + lambda I) ~ 1e7): single queries and small batches through the explicit inverse stay within
    1e-6 of the oracle (1e-3 on q-error)."""
    hl = lib.Handle(latency_mode=True)
    hl.fit(forest["x_train"], forest["y_train"])
    ref = oracle.Fit(forest["x_train"], forest["y_train"])
    rm, rv = ref.predict(forest["x_test"][:1024])
    for lo, hi in ((0, 1), (1, 9), (9, 1024)):
        m, v = hl.predict(forest["x_test"][lo:hi])
        assert relmax(m, rm[lo:hi]) < 1e-6
        assert np.max(np.abs(v - rv[lo:hi]) / np.abs(rv[lo:hi])) < 1e-6
        assert np.max(np.abs(2.0 ** np.abs(m - rm[lo:hi]) - 1.0)) < 1e-3


def test_latency_mode_with_a_tiny_regulariser(lib, synth):
    """diag_reg 1e-3 -> 1e-8 (cond up to ~1e11): the inverse-based variance still tracks LAPACK's substitution."""
    xtr, ytr, xte, _ = synth.make_problem(2000, 64, 32)
    for reg in (1e-5, 1e-8):
        hl = lib.Handle(diag_reg=reg, latency_mode=True)
        hl.fit(xtr, ytr)
        ref = oracle.Fit(xtr, ytr, diag_reg=reg)
        _, v = hl.predict(xte)
        _, rv = ref.predict(xte)
        kss = oracle.final_diag(oracle.layer0_diag(xte))
        # the variance is a cancellation kss - |v|^2: compare the subtracted norm, relative to kss
        assert np.max(np.abs(v - rv) / kss) < 1e-9

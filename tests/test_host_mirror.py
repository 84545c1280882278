"""CPU: the Python mirror of the neural-tangents surface (host logic only -- no arithmetic here).
The engine is replaced by an oracle-backed fake through runtime.new_handle, so laziness, caching, shapes,
error conventions and the lazy covariance are checked without a GPU."""
import numpy as np
import pytest

import nngp_oracle as oracle
from nngp_b200 import batch, predict, runtime, stax
from nngp_b200.estimator import Estimator


class FakeHandle:
    created = 0
    fits = 0

    def __init__(self, spec, diag_reg, absolute, kernel_type="nngp"):
        FakeHandle.created += 1
        self.spec, self.diag_reg, self.absolute, self.fit_, self.kt = spec, diag_reg, absolute, None, kernel_type

    def kernel(self, x1, x2=None):
        return oracle.kernel_fn(x1, x2, self.spec.depth, self.spec.sigma_w, self.spec.sigma_b, get=self.kt)

    def fit(self, x, y):
        FakeHandle.fits += 1
        cls = oracle.FitNTK if self.kt == "ntk" else oracle.Fit
        self.x_, self.y_ = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64).reshape(-1)
        self.fit_ = cls(x, y, self.spec.depth, self.spec.sigma_w, self.spec.sigma_b, self.diag_reg, self.absolute)

    def active_select(self, x_pool, budget, biased_sample=False, seed=10):          # nngp_active_select
        mean, var = self.fit_.predict(x_pool, True)
        pick = oracle.active_sample if biased_sample else oracle.active_select
        return pick(mean[:, None], np.sqrt(var), budget, *([seed] if biased_sample else []))

    def append_fit(self, x_new, y_new):                                             # nngp_append_fit
        self.fit(np.vstack([self.x_, x_new]), np.concatenate([self.y_, np.asarray(y_new).reshape(-1)]))

    def predict(self, x, want_var=True):
        return self.fit_.predict(x, want_var)


@pytest.fixture()
def fake_engine(monkeypatch):
    FakeHandle.created = FakeHandle.fits = 0
    monkeypatch.setattr(runtime, "new_handle",
                        lambda spec, diag_reg=0.0, diag_reg_absolute=False, kernel_type="nngp": FakeHandle(
                            spec, diag_reg, diag_reg_absolute, kernel_type))
    return FakeHandle


def test_serial_collapses_to_kernel_spec():
    init_fn, apply_fn, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))      # train.py:161-164
    assert kernel_fn.spec == stax.KernelSpec(depth=2, sigma_w=1.0, sigma_b=0.0)
    _, _, k3 = stax.serial(stax.Dense(512, W_std=1.5, b_std=0.05), stax.Relu(), stax.Dense(512, W_std=1.5, b_std=0.05),
                           stax.Relu(), stax.Dense(1, W_std=1.5, b_std=0.05))                      # active_train.py:44-49
    assert k3.spec == stax.KernelSpec(depth=3, sigma_w=1.5, sigma_b=0.05)
    assert batch.batch(kernel_fn, device_count=0, batch_size=0) is kernel_fn                        # train.py:166-168
    with pytest.raises(NotImplementedError):
        init_fn(None, (1, 2))
    for bad in ([stax.Dense(1), stax.Dense(1)], [stax.Relu(), stax.Dense(1)], [stax.Dense(1), stax.Relu()]):
        with pytest.raises(NotImplementedError):
            stax.serial(*bad)
    # Dense layers that differ in W_std / b_std [nt allows it; the reference never does]: per-layer spec
    _, _, kmix = stax.serial(stax.Dense(1, W_std=2.0), stax.Relu(), stax.Dense(1, b_std=0.3))
    assert kmix.spec == stax.KernelSpec(depth=2, sigma_w=(2.0, 1.0), sigma_b=(0.0, 0.3))
    with pytest.raises(NotImplementedError):
        stax.Conv(3, (3, 3))
    with pytest.raises(NotImplementedError):
        stax.Dense(1, parameterization="standard")


def test_per_layer_sigmas_in_the_oracle():
    """Per-layer W_std / b_std: closed form for depth 2 (K = W1^2 * kappa(k0, q0, q0') + b1^2 with k0 = W0^2 x.x'/D + b0^2),
    the uniform case as a special case, and the diagonal recursions."""
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 10, (7, 6))
    w, b = (1.7, 0.6), (0.2, 0.4)
    k = oracle.kernel_fn(x, None, 2, w, b)
    k0 = w[0] ** 2 * (x @ x.T) / 6 + b[0] ** 2
    q0 = np.diag(k0)
    s = np.sqrt(np.maximum(np.outer(q0, q0) - k0 ** 2, 0))
    th = np.arctan2(s, k0)
    want = w[1] ** 2 * (s / (2 * np.pi) + (0.5 - th / (2 * np.pi)) * k0) + b[1] ** 2
    assert np.allclose(k, want, rtol=1e-13)
    assert np.allclose(np.diag(k), oracle.final_diag(oracle.layer0_diag(x, w, b), 2, w, b), rtol=1e-13)
    assert np.array_equal(oracle.kernel_fn(x, None, 3, (1.5, 1.5, 1.5), (0.05,) * 3), oracle.kernel_fn(x, None, 3, 1.5, 0.05))
    th3 = oracle.kernel_fn(x, None, 3, (1.0, 1.3, 0.8), (0.1, 0.0, 0.2), get="ntk")
    assert np.allclose(np.diag(th3), oracle.final_diag_ntk(oracle.layer0_diag(x, (1.0, 1.3, 0.8), (0.1, 0.0, 0.2)), 3,
                                                           (1.0, 1.3, 0.8), (0.1, 0.0, 0.2)), rtol=1e-7)   # (theta ~ 1e-8 on the diagonal: DESIGN 5.5)
    with pytest.raises(ValueError):
        oracle.kernel_fn(x, None, 3, (1.0, 1.0), 0.0)


def test_kernel_fn_get_semantics(fake_engine):
    _, _, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    x = np.random.default_rng(0).uniform(0, 10, (5, 4))
    assert np.allclose(kernel_fn(x, None, "nngp"), oracle.kernel_fn(x))
    assert kernel_fn(x, x[:2], get="nngp").shape == (5, 2)
    assert np.allclose(kernel_fn(x, None, "ntk"), oracle.kernel_fn(x, get="ntk"))
    with pytest.raises(NotImplementedError):
        kernel_fn(x, None, ("nngp", "ntk"))
    with pytest.raises(ValueError):
        kernel_fn(x[0], None, "nngp")


def test_predict_fn_is_lazy_cached_and_shaped_like_neural_tangents(fake_engine):
    rng = np.random.default_rng(1)
    x, y, xt = rng.uniform(0, 10, (30, 6)), rng.uniform(0, 8, (30, 1)), rng.uniform(0, 10, (7, 6))
    _, _, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    predict_fn = predict.gradient_descent_mse_ensemble(kernel_fn, x, y, diag_reg=1e-3)             # train.py:171-172
    assert fake_engine.fits == 0                                  # construction does no math (BASELINE.md note)
    pred_mean, pred_cov = predict_fn(x_test=xt, get="nngp", compute_cov=True)                      # train.py:157-158
    assert fake_engine.fits == 1
    predict_fn(x_test=xt, get="nngp", compute_cov=True)
    assert fake_engine.fits == 1                                  # cached
    assert pred_mean.shape == (7, 1) and pred_cov.shape == (7, 7)
    ref = oracle.Fit(x, y)
    rm, rv = ref.predict(xt)
    pred_std = np.sqrt(np.diag(pred_cov))                                                           # train.py:180
    assert np.allclose(pred_mean.ravel(), rm) and np.allclose(pred_std, np.sqrt(rv))
    assert np.allclose(pred_cov.diagonal(), rv) and np.isclose(np.trace(pred_cov), rv.sum())
    with pytest.raises(NotImplementedError):
        np.asarray(pred_cov)                                      # the T x T matrix does not exist
    m_only = predict_fn(x_test=xt, get="nngp", compute_cov=False)
    assert m_only.shape == (7, 1)
    on_train = predict_fn(get="nngp")                             # x_test=None -> training inputs
    assert on_train.shape == (30, 1)
    # a new closure is a new fit (ActiveLearner.py:69,76 relies on it)
    predict.gradient_descent_mse_ensemble(kernel_fn, x, y, diag_reg=1e-3)(x_test=xt, get="nngp")
    assert fake_engine.fits == 2
    for kw in ({"t": 1.0}, {"get": ("nngp", "ntk")}, {"foo": 1}):
        with pytest.raises(NotImplementedError):
            predict_fn(x_test=xt, **kw)
    fits = fake_engine.fits
    m_ntk, c_ntk = predict_fn(x_test=xt, get="ntk", compute_cov=True)      # train.py:254 --kernel_type ntk
    predict_fn(x_test=xt, get="ntk", compute_cov=True)
    assert fake_engine.fits == fits + 1                                    # cached per `get`
    rn = oracle.FitNTK(x, y)
    assert np.allclose(m_ntk.ravel(), rn.predict(xt)[0]) and np.allclose(np.diag(c_ntk), rn.predict(xt)[1])
    y1 = y.ravel()
    assert predict.gradient_descent_mse_ensemble(kernel_fn, x, y1, diag_reg=1e-3)(x_test=xt, get="nngp").shape == (7,)
    with pytest.raises(ValueError):
        predict.gradient_descent_mse_ensemble(kernel_fn, x, y[:-1], diag_reg=1e-3)
    with pytest.raises(NotImplementedError):
        predict.gradient_descent_mse_ensemble(kernel_fn, x, np.zeros((30, 2)), diag_reg=1e-3)


def test_estimator_mirror(fake_engine):
    class Enc:
        def parse_line_without_card_then_encode(self, line):
            return np.array([float(v) for v in line.split(",")])
    rng = np.random.default_rng(2)
    x, y = rng.uniform(0, 10, (25, 4)), rng.uniform(0, 8, (25, 1))
    est = Estimator("schema", "data", "queries", X_train=x, Y_train=y, nngp_encoder=Enc(), verbose=False)
    assert fake_engine.fits == 0
    est.load_model()                                                                                # estimator.py:37-40
    assert fake_engine.fits == 1
    lines = ["1,2,3,4", "4,3,2,1", "0,0,5,5"]
    mean, std = est.predict(lines)                                                                  # estimator.py:42-61
    rm, rv = oracle.Fit(x, y).predict(np.array([[1, 2, 3, 4], [4, 3, 2, 1], [0, 0, 5, 5.0]]))
    assert mean.shape == (3,) and std.shape == (3,)
    assert np.allclose(mean, rm) and np.allclose(std, np.sqrt(rv))
    loaded = Estimator("s", "d", "q", 64, False, 100.0, 1.0, loader=lambda *a: (x, y, Enc()), verbose=False)
    assert loaded.X_train.shape == (25, 4)
    with pytest.raises(ValueError):
        Estimator("s", "d", "q", verbose=False)


def test_active_learner_mirror_matches_oracle_selection(fake_engine):
    """Config C4 logic (active/ActiveLearner.py:43-77): deterministic top-k branch vs the oracle's rule."""
    from nngp_b200.active import ActiveLearner
    from nngp_b200 import synth
    xtr, ytr, xpool, ypool = synth.make_problem(60, 90, 8)
    xval, yval = synth.encodings(30, 8, 5), None
    yval = synth.labels(xval)
    _, _, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    al = ActiveLearner(budget=20, active_iters=2, verbose=False)
    pf = al.train(kernel_fn, xtr, ytr[:, None])
    idx = al.active_test(pf, xpool)
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xpool)
    assert list(idx) == list(oracle.active_select(rm[:, None], np.sqrt(rv), 20))
    fits0 = fake_engine.fits
    pf2, x_end, y_end = al.active_train(kernel_fn, xtr, ytr[:, None], xpool, ypool[:, None], xval, yval[:, None])
    assert x_end.shape == (60 + 2 * 20, 8) and y_end.shape == (100, 1)
    assert fake_engine.fits - fits0 == 3 and len(al.history) == 3      # refit from scratch every iteration
    x2, y2, xp2, yp2 = al.merge_data(idx, xtr, ytr[:, None], xpool, ypool[:, None])
    assert x2.shape[0] == 80 and xp2.shape[0] == 70 and not set(map(tuple, xp2)) & set(map(tuple, xpool[idx]))


def test_compat_shims_resolve_to_the_mirror():
    import importlib
    import sys
    from pathlib import Path
    compat = str(Path(__file__).resolve().parents[1] / "nngp-src_b200" / "compat")
    sys.path.insert(0, compat)
    try:
        nt = importlib.import_module("neural_tangents")
        jnp = importlib.import_module("jax.numpy")
        from jax.config import config
        from jax.lib import xla_bridge
        config.update("jax_enable_x64", True)
        assert nt.stax is stax and nt.predict is predict and nt.batch is batch.batch
        assert xla_bridge.get_backend().platform == "cpu"
        cov = predict.LazyCovariance(np.array([4.0, 9.0]))
        assert np.allclose(jnp.sqrt(jnp.diag(cov)), [2.0, 3.0])
    finally:
        sys.path.remove(compat)
        for m in [k for k in sys.modules if k == "jax" or k.startswith("jax.") or k == "neural_tangents"]:
            del sys.modules[m]


REFERENCE_AL = "/root/reference/active/ActiveLearner.py"


@pytest.mark.skipif(not __import__("os").path.exists(REFERENCE_AL),
                    reason="the reference tree is only mounted in the build container")
def test_unmodified_reference_active_learner_runs_on_the_shims(fake_engine):
    """Drop-in check for row a8: the reference's OWN active/ActiveLearner.py, imported unmodified from the reference
    tree with nngp-src_b200/compat on sys.path (jax / neural_tangents shims), drives the mirror -- train, test,
    active_test (np.sqrt(np.diag(pred_cov)) on the lazy covariance, argsort tail), merge_data, refit -- and ends with
    the same training set as nngp_b200.active.ActiveLearner.  Only `util` (seaborn / matplotlib plotting helpers,
    out of scope) is stubbed."""
    import importlib.util
    import sys
    import types
    from pathlib import Path
    from nngp_b200 import synth
    from nngp_b200.active import ActiveLearner as MirrorLearner
    compat = str(Path(__file__).resolve().parents[1] / "nngp-src_b200" / "compat")
    util_stub = types.ModuleType("util")

    class PredictionStatistics:                      # util.py:152-167 prints q-error tables; irrelevant here
        def get_prediction_details(self, *a, **k):
            return None

    util_stub.PredictionStatistics = PredictionStatistics
    sys.path.insert(0, compat)
    sys.modules["util"] = util_stub
    try:
        spec = importlib.util.spec_from_file_location("reference_active_learner", REFERENCE_AL)
        ref_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_mod)
        nt = importlib.import_module("neural_tangents")
        _, _, kernel_fn = nt.stax.serial(nt.stax.Dense(512), nt.stax.Relu(), nt.stax.Dense(1))
        xtr, ytr, xpool, ypool = synth.make_problem(50, 80, 8)
        xval, yval = synth.encodings(20, 8, 5), None
        yval = synth.labels(xval)
        args = types.SimpleNamespace(budget=15, active_iters=2, kernel_type="nngp", biased_sample=False)
        learner = ref_mod.ActiveLearner(args)
        learner.active_train(kernel_fn, xtr, ytr[:, None], xpool, ypool[:, None], xval, yval[:, None], None)
        # the reference's loop keeps its arrays local; replay its selection rule through its own methods
        pf = learner.train(kernel_fn, xtr, ytr[:, None])
        sel = np.asarray(learner.active_test(pf, xpool, "nngp"))
        x1, y1, xp1, yp1 = learner.merge_data(sel, xtr, ytr[:, None], xpool, ypool[:, None])
        mirror = MirrorLearner(budget=15, active_iters=0, verbose=False)
        pf_m = mirror.train(kernel_fn, xtr, ytr[:, None])
        sel_m = np.asarray(mirror.active_test(pf_m, xpool))
        assert sel.shape == (15,) and list(sel) == list(sel_m)
        assert x1.shape == (65, 8) and xp1.shape == (65, 8) and y1.shape == (65, 1)
        x1m, _, xp1m, _ = mirror.merge_data(sel_m, xtr, ytr[:, None], xpool, ypool[:, None])
        assert np.array_equal(np.asarray(x1), x1m) and np.array_equal(np.asarray(xp1), xp1m)
        assert fake_engine.fits >= 4               # 3 fits inside active_train + the replay
    finally:
        sys.path.remove(compat)
        sys.modules.pop("util", None)
        for m in [k for k in sys.modules if k == "jax" or k.startswith("jax.") or k == "neural_tangents"]:
            del sys.modules[m]


REFERENCE_TRAIN = "/root/reference/train.py"


@pytest.mark.skipif(not __import__("os").path.exists(REFERENCE_TRAIN),
                    reason="the reference tree is only mounted in the build container")
def test_unmodified_reference_train_py_nngp_path_runs_on_the_shims(fake_engine, capsys):
    """Drop-in check for rows a1-a6 at their primary call site: the reference's OWN train.py, imported unmodified with
    the shims on sys.path, runs NNGP_train_and_test (train.py:153-203: stax.serial -> nt.batch ->
    gradient_descent_mse_ensemble -> predict_fn(get='nngp', compute_cov=True) -> sqrt(diag(cov))) and prints the
    squared error the oracle gives.  Stubbed: `datasets`, `schemas` (DB loaders) and `util` (plotting / statistics)."""
    import importlib.util
    import re
    import sys
    import types
    from pathlib import Path
    from nngp_b200 import synth
    compat = str(Path(__file__).resolve().parents[1] / "nngp-src_b200" / "compat")
    util_stub = types.ModuleType("util")

    class PredictionStatistics:
        def get_prediction_details(self, *a, **k):
            return None

    util_stub.PredictionStatistics = PredictionStatistics
    for name in ("draw_uncertainty", "calibration_plot", "draw_kernel_heatmap", "show_memory_usage",
                 "uneven_train_test_split", "train_test_val_split"):
        setattr(util_stub, name, lambda *a, **k: None)
    stubs = {"util": util_stub, "datasets": types.ModuleType("datasets"), "schemas": types.ModuleType("schemas")}
    sys.path.insert(0, compat)
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("reference_train", REFERENCE_TRAIN)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        xtr, ytr, xte, yte = synth.make_problem(70, 25, 10)
        args = types.SimpleNamespace(kernel_type="nngp", cuda=False)
        mod.NNGP_train_and_test(args, xtr, ytr[:, None], xte, yte[:, None], None, None)
        out = capsys.readouterr().out
        printed = float(re.search(r"Mean Square Error: ([-+0-9.eE]+)", out).group(1))
        rm, _ = oracle.Fit(xtr, ytr).predict(xte)
        assert abs(printed - float(np.sum((rm - yte) ** 2))) <= 1e-9 * max(1.0, abs(printed))
        assert "Kernel construction in" in out and "Inference time=" in out
        assert fake_engine.fits == 1               # lazy fit on the first predict_fn call, cached for the second
        # train.py:254 --kernel_type ntk goes through the same function
        mod.NNGP_train_and_test(types.SimpleNamespace(kernel_type="ntk", cuda=False), xtr, ytr[:, None], xte, yte[:, None])
        printed = float(re.search(r"Mean Square Error: ([-+0-9.eE]+)", capsys.readouterr().out).group(1))
        rm_ntk, _ = oracle.FitNTK(xtr, ytr).predict(xte)
        assert abs(printed - float(np.sum((rm_ntk - yte) ** 2))) <= 1e-9 * max(1.0, abs(printed))
    finally:
        sys.path.remove(compat)
        for k in stubs:
            sys.modules.pop(k, None)
        for m in [k for k in sys.modules if k == "jax" or k.startswith("jax.") or k == "neural_tangents"]:
            del sys.modules[m]


REFERENCE_ACTIVE_TRAIN = "/root/reference/active/active_train.py"


@pytest.mark.skipif(not __import__("os").path.exists(REFERENCE_ACTIVE_TRAIN),
                    reason="the reference tree is only mounted in the build container")
def test_unmodified_reference_active_train_driver_runs_on_the_shims(fake_engine, capsys):
    """The reference's OWN active/active_train.py main() (config C4's driver, active_train.py:21-51) with its default
    biased_sample=True branch (jax.random.choice -> the shim's weighted draw): loaders stubbed, everything else as
    shipped -- including `from active.ActiveLearner import ActiveLearner` resolved from the reference tree."""
    import importlib.util
    import sys
    import types
    from pathlib import Path
    from nngp_b200 import synth
    compat = str(Path(__file__).resolve().parents[1] / "nngp-src_b200" / "compat")
    x = synth.encodings(200, 8, 1)
    y = synth.labels(x)[:, None]
    util_stub = types.ModuleType("util")

    class PredictionStatistics:
        def get_prediction_details(self, *a, **k):
            return None

    def train_test_val_split(X, Y, train_frac=0.2, test_frac=0.6, all_query_infos=None):       # util.py:271-293
        n = X.shape[0]
        a, b = int(n * train_frac), int(n * (train_frac + test_frac))
        return X[:a], Y[:a], None, X[a:b], Y[a:b], None, X[b:], Y[b:], None

    util_stub.PredictionStatistics = PredictionStatistics
    util_stub.train_test_val_split = train_test_val_split
    schemas_stub = types.ModuleType("schemas")
    schemas_stub.load_training_schema_data = lambda args: (x, y, None)
    stubs = {"util": util_stub, "datasets": types.ModuleType("datasets"), "schemas": schemas_stub}
    sys.path[:0] = [compat, "/root/reference"]
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("reference_active_train", REFERENCE_ACTIVE_TRAIN)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        args = types.SimpleNamespace(kernel_type="nngp", biased_sample=True, active_iters=3, budget=25)
        mod.main(args)
        out = capsys.readouterr().out
        assert "# Initial Training samples: 40" in out
        grown = [l for l in out.splitlines() if l.startswith("# Training samples:")]
        assert grown == ["# Training samples: 65", "# Training samples: 90", "# Training samples: 115"]
        assert out.count("Test MSE Loss:") == 4 and fake_engine.fits == 4
    finally:
        for p in (compat, "/root/reference"):
            sys.path.remove(p)
        for k in stubs:
            sys.modules.pop(k, None)
        for m in [k for k in sys.modules if k == "jax" or k.startswith("jax.") or k == "neural_tangents"
                  or k == "active" or k.startswith("active.")]:
            del sys.modules[m]


def test_first_party_cli_loader_reproduces_the_forest_fixture(forest, fake_engine, capsys, tmp_path):
    """python -m nngp_b200.train (mirror of train.py:225-298): its loader + split on the reference's shipped forest
    queries give exactly the committed C1 fixture (made by tests/golden/make_forest_fixture.py); on boxes without the
    reference tree a small synthetic query directory exercises the same code.  The run prints train.py's lines and the
    oracle's squared error (the engine is the oracle-backed fake here; the GPU test drives the real one)."""
    import os
    import re
    from nngp_b200 import train as cli
    qdir = "/root/reference/Queries/forest_data"
    have_ref = os.path.isdir(qdir)
    if not have_ref:
        qdir = str(tmp_path)
        rng = np.random.default_rng(0)
        with open(tmp_path / "query_1.txt", "w") as fh:
            for _ in range(200):
                preds = []
                for c, (lo, hi) in cli.FOREST_RANGES.items():
                    if rng.random() < 0.5:
                        a, b = sorted(rng.integers(lo, hi + 1, 2).tolist())
                        preds.append(f"{c},{b},{a}")
                fh.write("#".join(preds or ["A,3000,2000"]) + f"@{int(rng.integers(1, 5000))}\n")
    args = cli.build_parser().parse_args(["--query_path", qdir])
    args.join_query = False
    x, y = cli.load_training_data(args)
    xtr, ytr, xte, yte, _, _ = cli.train_test_val_split(x, y)
    if have_ref:
        assert np.array_equal(xtr, forest["x_train"]) and np.array_equal(ytr[:, 0], forest["y_train"])
        assert np.array_equal(xte, forest["x_test"]) and np.array_equal(yte[:, 0], forest["y_test"])
        xtr, ytr, xte, yte = xtr[:300], ytr[:300], xte[:100], yte[:100]        # keep the CPU fake quick
    capsys.readouterr()
    _, std, mse = cli.NNGP_train_and_test(args, xtr, ytr, xte, yte)
    out = capsys.readouterr().out
    rm, rv = oracle.Fit(xtr, ytr).predict(xte)
    assert abs(mse - float(np.sum((rm[:, None] - yte) ** 2))) <= 1e-9 * max(1.0, mse)
    assert np.allclose(std, np.sqrt(rv), rtol=1e-9)
    assert float(re.search(r"Mean Square Error: ([-+0-9.eE]+)", out).group(1)) == mse
    assert "Kernel construction in" in out and "Inference time=" in out and "q-error: median" in out
    with pytest.raises(NotImplementedError):
        a2 = cli.build_parser().parse_args(["--relations", "a,b"])
        a2.join_query = True
        cli.main(a2)


def test_retrain_takes_the_engine_and_the_old_closure_stays_valid(fake_engine):
    """ActiveLearner.retrain appends to the handle of the previous predict_fn.  That closure must not start
    answering with the new model: it gives the engine up and refits from its own data when used again (in the
    reference every gradient_descent_mse_ensemble call is an independent fit, active/ActiveLearner.py:69,76)."""
    from nngp_b200 import synth
    from nngp_b200.active import ActiveLearner
    _, _, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    xtr, ytr, xpool, ypool = synth.make_problem(60, 40, 8)
    al = ActiveLearner(budget=10, active_iters=1, verbose=False)
    old_fn = al.train(kernel_fn, xtr, ytr[:, None])
    before = old_fn(x_test=xpool, get="nngp", compute_cov=False)
    sel = al.active_test(old_fn, xpool)
    x2, y2, _, _ = al.merge_data(sel, xtr, ytr[:, None], xpool, ypool[:, None])
    new_fn = al.retrain(kernel_fn, old_fn, x2, y2, xpool[sel], ypool[sel][:, None])
    after_new = new_fn(x_test=xpool, get="nngp", compute_cov=False)
    after_old = old_fn(x_test=xpool, get="nngp", compute_cov=False)
    assert not np.allclose(after_new, before)                 # the new closure serves the extended model
    assert np.array_equal(after_old, before)                  # the old one still serves the old model (refitted)
    assert new_fn.engine() is not old_fn.engine()

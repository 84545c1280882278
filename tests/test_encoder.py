"""CPU: the C++ batch query-line encoder (scope row f-1) against vectors produced by RUNNING THE REFERENCE'S OWN
encoder (tests/golden/make_encoder_golden.py -> encoder_golden.npz): bit-exact float64 rows."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def gold(golden_dir):
    z = np.load(golden_dir / "encoder_golden.npz")
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def encmod():
    import __graft_entry__ as g
    g.build()
    from nngp_b200 import encoder
    return encoder


def test_multi_table_lines_bit_exact(encmod, gold):
    enc = encmod.BatchEncoder(str(gold["schema"]))
    assert enc.dim == gold["x"].shape[1]
    lines = [str(l) for l in gold["lines"]]
    x = enc.encode(lines)
    assert x.dtype == np.float64 and np.array_equal(x, gold["x"])            # bit for bit, 600 lines
    assert np.array_equal(encmod.BatchEncoder(str(gold["schema"]), n_threads=1).encode(lines), gold["x"])
    assert np.array_equal(enc.parse_line_without_card_then_encode(lines[3]), gold["x"][3])
    big = enc.encode(lines * 5)                                               # threaded path (>= 256 lines)
    assert np.array_equal(big, np.tile(gold["x"], (5, 1)))


def test_training_lines_with_cardinality(encmod, gold):
    enc = encmod.BatchEncoder(str(gold["schema"]))
    x, card = enc.encode([str(l) for l in gold["train_lines"]], fmt=encmod.FORMAT_TRAIN_LINE, with_card=True)
    assert np.array_equal(x, gold["x_train"]) and np.array_equal(card, gold["cards"])


def test_single_table_forest_lines(encmod, gold, forest):
    enc = encmod.BatchEncoder(str(gold["forest_schema"]))
    x, card = enc.encode([str(l) for l in gold["forest_lines"]], fmt=encmod.FORMAT_SINGLE_TABLE, with_card=True)
    assert np.array_equal(x, gold["forest_x"]) and np.array_equal(card, gold["forest_cards"])
    # the fixture generator's restatement of the same encoding (tests/golden/make_forest_fixture.py) agrees too:
    # every encoded golden row appears in the shuffled forest fixture
    rows = {r.tobytes() for r in np.vstack([forest["x_train"], forest["x_test"]])}
    assert sum(r.tobytes() in rows for r in x) >= int(0.75 * len(x))          # 60/20 of the 60/20/20 split are stored


def test_errors_are_reported_with_line_numbers(encmod, gold):
    enc = encmod.BatchEncoder(str(gold["schema"]))
    good = str(gold["lines"][0])
    for bad, msg in [("nosuch@@", "unknown table"), ("orders@nosuch,1,0@", "unknown column"),
                     ("orders@o_total,abc,0@", "bad numerical"), ("orders,customer@@", "Format"),
                     ("orders,customer@@@orders,customer,o_key", "join")]:
        with pytest.raises(ValueError, match=msg):
            enc.encode([good, bad])
    with pytest.raises(ValueError, match="line 1"):
        enc.encode([good, "nosuch@@"])
    with pytest.raises(ValueError):
        encmod.BatchEncoder("table t\ncol a weird 1 2")
    with pytest.raises(ValueError):
        encmod.BatchEncoder("chunk_size 128\ntable t\ncol a num 0 1")
    assert enc.encode([]).shape == (0, enc.dim)


def test_estimator_uses_the_batch_encoder(encmod, gold, monkeypatch):
    import nngp_oracle as oracle
    from nngp_b200 import runtime
    from nngp_b200.estimator import Estimator

    class FakeHandle:
        def __init__(self, spec, diag_reg, absolute):
            self.spec, self.diag_reg = spec, diag_reg

        def fit(self, x, y):
            self.f = oracle.Fit(x, y, self.spec.depth, diag_reg=self.diag_reg)

        def predict(self, x, want_var=True):
            return self.f.predict(x, want_var)

    monkeypatch.setattr(runtime, "new_handle", lambda spec, diag_reg=0.0, diag_reg_absolute=False, kernel_type="nngp": FakeHandle(spec, diag_reg, diag_reg_absolute))
    enc = encmod.BatchEncoder(str(gold["schema"]))
    y = np.log2(gold["cards"])[:, None]
    est = Estimator("s", "d", "q", X_train=gold["x_train"], Y_train=y, nngp_encoder=enc, verbose=False)
    mean, std = est.predict([str(l) for l in gold["lines"][:20]])
    rm, rv = oracle.Fit(gold["x_train"], y).predict(gold["x"][:20])
    assert np.allclose(mean, rm) and np.allclose(std, np.sqrt(rv))
    # large batches are pipelined (encode chunk i+1 on the host while chunk i is predicted): same answer, same order
    lines = [str(l) for l in gold["lines"][:20]] * 3 + [str(gold["lines"][3])]
    est.pipeline_chunk = 7
    mean_p, std_p = est.predict(lines)
    assert mean_p.shape == (61,) and std_p.shape == (61,)
    assert np.allclose(mean_p[:60], np.tile(rm, 3), rtol=1e-12) and np.allclose(std_p[:60], np.tile(np.sqrt(rv), 3), rtol=1e-9)
    assert np.isclose(mean_p[60], rm[3], rtol=1e-12)


REFERENCE_ESTIMATOR = "/root/reference/neuroestimator/estimator/estimator.py"


@pytest.mark.skipif(not __import__("os").path.exists(REFERENCE_ESTIMATOR),
                    reason="the reference tree is only mounted in the build container")
def test_unmodified_reference_estimator_runs_on_the_shims(encmod, gold, monkeypatch):
    """Drop-in check for row a7: the reference's OWN neuroestimator/estimator/estimator.py, imported unmodified with
    nngp-src_b200/compat on sys.path, builds its kernel through the stax shim, fits lazily in load_model() and serves
    predict(query_lines) -> (mean, std) through the mirror; `estimator.util.load_training_schema_data` (pandas schema
    loading, out of scope) is stubbed to hand over the fixture's training set and the C++ batch encoder, whose
    parse_line_without_card_then_encode the reference's per-line loop calls."""
    import importlib.util
    import sys
    import types
    from pathlib import Path
    import nngp_oracle as oracle
    from nngp_b200 import runtime

    class FakeHandle:
        def __init__(self, spec, diag_reg, absolute):
            self.spec, self.diag_reg = spec, diag_reg

        def fit(self, x, y):
            self.f = oracle.Fit(x, y, self.spec.depth, diag_reg=self.diag_reg)

        def predict(self, x, want_var=True):
            return self.f.predict(x, want_var)

    monkeypatch.setattr(runtime, "new_handle", lambda spec, diag_reg=0.0, diag_reg_absolute=False, kernel_type="nngp": FakeHandle(spec, diag_reg, diag_reg_absolute))
    enc = encmod.BatchEncoder(str(gold["schema"]))
    y = np.log2(gold["cards"])[:, None]
    compat = str(Path(__file__).resolve().parents[1] / "nngp-src_b200" / "compat")
    pkg = types.ModuleType("estimator")
    pkg.__path__ = []
    util = types.ModuleType("estimator.util")
    util.load_training_schema_data = lambda *a, **k: (gold["x_train"], y, enc)
    sys.path.insert(0, compat)
    sys.modules["estimator"], sys.modules["estimator.util"] = pkg, util
    try:
        spec = importlib.util.spec_from_file_location("estimator.estimator", REFERENCE_ESTIMATOR)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["estimator.estimator"] = mod
        spec.loader.exec_module(mod)
        est = mod.Estimator("tpch", "/data", "/queries")
        est.load_model()
        mean, std = est.predict([str(l) for l in gold["lines"][:20]])
        rm, rv = oracle.Fit(gold["x_train"], y).predict(gold["x"][:20])
        assert mean.shape == (20,) and std.shape == (20,)
        assert np.allclose(mean, rm) and np.allclose(std, np.sqrt(rv))
    finally:
        sys.path.remove(compat)
        for m in [k for k in sys.modules if k == "jax" or k.startswith("jax.") or k == "neural_tangents"
                  or k == "estimator" or k.startswith("estimator.")]:
            del sys.modules[m]

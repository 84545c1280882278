"""GPU tests at BASELINE.json's full sizes (C3: N = 32 768, D = 256, depth 3; C5: N = 16 384, D = 512, depth 3, 96 join
dims), the Estimator / train.py call sequences on the real engine (rows a7, a1-a6), the in-process multi-GPU handle
and the packed state path.  Everything goes through the C ABI; the oracle (numpy/scipy, all host cores) is the checker.

Tolerances (north_star): predicted log-cardinalities within 1e-6 relative in FP64, 1e-3 on q-error.
"""
import json
import os
import re
import subprocess
import sys
import time
import types
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import nngp_oracle as oracle  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]


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


def _record(name, rec):
    """Parity numbers of the full-size runs -> gpurun_out/ (copied into profiles/ by the builder)."""
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        (out / f"parity_{name}.json").write_text(json.dumps(rec, indent=1))
    except OSError:
        pass


def _full_size_case(lib, synth, name, n, t, d, depth, join_dims, sample, slices=0):
    """Fit at the config's full N, predict `t` test rows (the bench's per-GPU batch), check
       (i)   lambda = 1e-3 * mean(q_L) from the COMPUTED diagonal,
       (ii)  the posterior at the training rows: mean = y - lambda*alpha, 0 < var < lambda,
       (iii) a sampled-row ORACLE check (`sample` rows spread over the batch) at 1e-6 (1e-3 on q-error),
       (iv)  the sampled rows predicted alone (few row tiles: the pipelined kernel variant) == the same rows inside the
             big batch, bit for bit,
       (v)   with `slices`: the same batch on a handle with variance_slices (variance product on the int8 tensor
             cores): mean bitwise the FP64 handle's, sampled variances against the oracle at the same 1e-6."""
    xtr = synth.encodings(n, d, 1, join_dims=join_dims)
    ytr = synth.labels(xtr, join_dims)
    if np.ptp(ytr) < 1.0:                       # the independence surrogate saturates at wide encodings
        ytr = np.random.default_rng(3).uniform(0.0, 20.0, n)
    xte = synth.encodings(t, d, 2, join_dims=join_dims)
    h = lib.Handle(depth=depth, stats_level=1)
    t0 = time.perf_counter()
    h.fit(xtr, ytr)
    fit_s = time.perf_counter() - t0
    nn, dd, lam = h.dims()
    q0 = np.einsum("ij,ij->i", xtr, xtr) / d
    assert (nn, dd) == (n, d)
    assert abs(lam - 1e-3 * np.mean(q0) / 2 ** (depth - 1)) < 1e-12 * lam
    alpha = h.get_state(x=False, l=False)["alpha"]
    mean_tr, var_tr = h.predict(xtr)
    assert np.max(np.abs(mean_tr - (ytr - lam * alpha))) < 1e-6 * np.max(np.abs(ytr))
    assert np.all(var_tr > 0) and np.all(var_tr < lam * (1 + 1e-9))
    t0 = time.perf_counter()
    mean, var = h.predict(xte)
    pred_s = time.perf_counter() - t0
    assert np.all(np.isfinite(mean)) and np.all(var > 0)
    idx = np.linspace(0, t - 1, sample).astype(np.int64)
    m_s, v_s = h.predict(xte[idx])
    assert np.array_equal(m_s, mean[idx]) and np.array_equal(v_s, var[idx])
    stats = h.stats()
    h.close()
    sliced = None
    if slices:
        hs = lib.Handle(depth=depth, stats_level=1, variance_slices=slices)
        hs.fit(xtr, ytr)
        t0 = time.perf_counter()
        mean_s, var_s = hs.predict(xte)
        sliced = {"slices": slices, "predict_wall_s": time.perf_counter() - t0, "sliced_ms": hs.stats()["sliced_ms"],
                  "var_max_rel_vs_fp64_path": float(np.max(np.abs(var_s - var) / np.abs(var)))}
        assert np.array_equal(mean_s, mean) and np.all(var_s > 0)
        hs.close()
    t0 = time.perf_counter()
    ref = oracle.Fit(xtr, ytr, depth)
    ofit_s = time.perf_counter() - t0
    assert abs(lam - ref.lam) < 1e-12 * ref.lam
    assert relmax(alpha, ref.alpha) < 1e-6
    rm, rv = ref.predict(xte[idx])
    e_mean, e_var = relmax(m_s, rm), float(np.max(np.abs(v_s - rv) / np.abs(rv)))
    e_q = float(np.max(np.abs(2.0 ** np.abs(m_s - rm) - 1.0)))
    if sliced:
        sliced["max_rel_err_var_vs_oracle"] = float(np.max(np.abs(var_s[idx] - rv) / np.abs(rv)))
        sliced["max_rel_err_std_vs_oracle"] = float(np.max(np.abs(np.sqrt(var_s[idx]) - np.sqrt(rv)) / np.sqrt(rv)))
    _record(name, {"sliced": sliced, "config": {"n_train": n, "test_rows": t, "dim": d, "depth": depth, "join_dims": join_dims},
                   "sampled_rows": int(sample), "max_rel_err_mean": e_mean, "max_rel_err_var": e_var,
                   "max_q_error_dev": e_q, "alpha_rel_err": relmax(alpha, ref.alpha), "lambda": lam,
                   "oracle_lambda": ref.lam, "gpu_fit_wall_s": fit_s, "gpu_fit_device_s": stats["fit_total_ms"] / 1e3,
                   "gpu_predict_wall_s": pred_s, "oracle_fit_s": ofit_s, "host_cores": len(os.sched_getaffinity(0)),
                   "tolerance": "1e-6 relative on mean / variance, 1e-3 on q-error (north_star)"})
    assert e_mean < 1e-6 and e_var < 1e-6 and e_q < 1e-3
    if sliced:
        assert sliced["max_rel_err_var_vs_oracle"] < 1e-6, sliced


def test_c3_full_size_parity(lib, synth):
    """BASELINE config C3 per GPU: N = 32 768 (W = 512 Cholesky panels), D = 256, depth 3, 131 072 test rows (one
    34 GB row block).  The oracle's fit at this size is ~1-2 minutes of host CPU."""
    _full_size_case(lib, synth, "c3", 32768, 131072, 256, 3, 0, 512, slices=7)


def test_c5_full_size_parity(lib, synth):
    """BASELINE config C5: N = 16 384, D = 512 with 96 join dims (JoinQuerySampler layout), depth 3; 262 144 test rows."""
    _full_size_case(lib, synth, "c5", 16384, 262144, 512, 3, 96, 512)


# ---------------------------------------------------------------------------------------------- a7: Estimator
def test_estimator_end_to_end_on_the_engine(lib, golden_dir):
    """Row a7 on hardware: query LINES -> C++ batch encoder -> Estimator.load_model / predict on the real engine ->
    (mean, std); checked against the oracle on the golden encodings produced by the reference's own encoder classes
    (neuroestimator/estimator/estimator.py:16-68)."""
    from nngp_b200 import encoder as encmod
    from nngp_b200.estimator import Estimator
    gold = np.load(golden_dir / "encoder_golden.npz", allow_pickle=True)
    enc = encmod.BatchEncoder(str(gold["schema"]))
    y = np.log2(gold["cards"])[:, None]
    est = Estimator("s", "d", "q", X_train=gold["x_train"], Y_train=y, nngp_encoder=enc, verbose=False)
    est.load_model()
    lines = [str(l) for l in gold["lines"]]
    mean, std = est.predict(lines)
    ref = oracle.Fit(gold["x_train"], y)
    rm, rv = ref.predict(gold["x"])
    assert mean.shape == (len(lines),) and std.shape == (len(lines),)
    assert relmax(mean, rm) < 1e-6 and relmax(std, np.sqrt(rv)) < 1e-6
    # the pipelined path (encode chunk i+1 on the host while the GPU predicts chunk i) returns the same bits
    est.pipeline_chunk = 97
    mean_p, std_p = est.predict(lines)
    assert np.array_equal(mean_p, mean) and np.array_equal(std_p, std)
    # the handle behind predict_fn is the real library
    assert isinstance(est.predict_fn.engine(), lib.Handle)


REFERENCE_TRAIN = "/root/reference/train.py"


def test_train_py_call_sequence_on_the_engine(lib, forest, capsys):
    """Rows a1-a6 at their primary call site, on hardware and on config C1 (the shipped forest queries): the exact
    call sequence of train.py:153-203 -- stax.serial -> nt.batch -> gradient_descent_mse_ensemble ->
    predict_fn(get='nngp', compute_cov=True) -> sqrt(diag(cov)) -> MSE print.  When the reference tree is mounted
    (build container) its OWN train.py runs unmodified on the import shims; on the GPU box the same sequence runs
    through nngp_b200.nt.  The printed squared error and the q-error summary must be the oracle's."""
    xtr, ytr, xte, yte = forest["x_train"], forest["y_train"][:, None], forest["x_test"], forest["y_test"][:, None]
    if os.path.exists(REFERENCE_TRAIN):
        import importlib.util
        compat = str(ROOT / "nngp-src_b200" / "compat")
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
            mod.NNGP_train_and_test(types.SimpleNamespace(kernel_type="nngp", cuda=False), xtr, ytr, xte, yte, None, None)
        finally:
            sys.path.remove(compat)
            for k in stubs:
                sys.modules.pop(k, None)
            for m in [k for k in sys.modules if k == "jax" or k.startswith("jax.") or k == "neural_tangents"]:
                del sys.modules[m]
        out = capsys.readouterr().out
        printed = float(re.search(r"Mean Square Error: ([-+0-9.eE]+)", out).group(1))
    else:
        printed = None
    # the same sequence through the first-party names
    import nngp_b200.nt as nt
    from nngp_b200 import stax
    _, _, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    kernel_fn = nt.batch(kernel_fn, device_count=0, batch_size=0)
    predict_fn = nt.predict.gradient_descent_mse_ensemble(kernel_fn, xtr, ytr, diag_reg=1e-3)
    pred_mean, pred_cov = predict_fn(x_test=xte, get="nngp", compute_cov=True)
    pred_std = np.sqrt(np.diag(pred_cov))
    mse = float(np.sum(np.power(pred_mean - yte, 2)))
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xte)
    want = float(np.sum((rm[:, None] - yte) ** 2))
    assert abs(mse - want) <= 1e-6 * want
    if printed is not None:
        assert abs(printed - want) <= 1e-6 * want and printed == mse       # the shims and the first-party path: same bits
    assert pred_mean.shape == (xte.shape[0], 1) and pred_std.shape == (xte.shape[0],)
    assert relmax(pred_std, np.sqrt(rv)) < 1e-6
    q, rq = oracle.q_error_stats(pred_mean, yte), oracle.q_error_stats(rm, yte)
    for k in ("median", "mean", "p95", "max"):
        assert abs(q[k] - rq[k]) <= 1e-3 * rq[k]
    assert isinstance(predict_fn.engine(), lib.Handle)


# ---------------------------------------------------------------------------------------------- packed state
def test_packed_state_roundtrip(lib, synth):
    """nngp_state_pack / _unpack (what broadcast_fit ships between processes): chunked, host and device buffers."""
    import torch
    xtr, ytr, xte, _ = synth.make_problem(777, 400, 18)
    h = lib.Handle()
    h.fit(xtr, ytr)
    mean, var = h.predict(xte)
    n, d, lam = h.dims()
    total = h.packed_size()
    assert total == n * d + n + n * (n + 1) // 2
    flat = np.empty(total)
    h.state_pack(0, total, flat)
    st = h.get_state()
    assert np.array_equal(flat[:n * d].reshape(n, d), st["x"]) and np.array_equal(flat[n * d:n * d + n], st["alpha"])
    assert np.array_equal(flat[n * d + n:], st["l"][np.tril_indices(n)])
    h2 = lib.Handle()
    h2.state_import_begin(n, d)
    dev = torch.empty(100_000, dtype=torch.float64, device="cuda")
    for off in range(0, total, 100_000):                      # device staging chunks, as under NCCL
        cnt = min(100_000, total - off)
        h.state_pack(off, cnt, dev[:cnt])
        h2.state_unpack(off, cnt, dev[:cnt])
    h2.state_import_end(lam)
    m2, v2 = h2.predict(xte)
    assert np.array_equal(m2, mean) and np.array_equal(v2, var)
    h3 = lib.Handle()
    h3.state_import_begin(n, d)
    h3.state_unpack(0, total, flat)                           # host buffer
    h3.state_import_end(lam)
    m3, v3 = h3.predict(xte)
    assert np.array_equal(m3, mean) and np.array_equal(v3, var)
    with pytest.raises(ValueError):
        h.state_pack(total - 5, 10, np.empty(10))
    with pytest.raises(lib.NngpError):
        lib.Handle().state_unpack(0, 4, np.zeros(4))


def test_output_buffers_must_be_float64_contiguous(lib, synth):
    """a float32 / strided output buffer used to be converted into a temporary and come back unfilled."""
    xtr, ytr, xte, _ = synth.make_problem(200, 50, 8)
    h = lib.Handle()
    h.fit(xtr, ytr)
    with pytest.raises(ValueError):
        h.predict(xte, mean_out=np.empty(50, dtype=np.float32))
    with pytest.raises(ValueError):
        h.predict(xte, mean_out=np.empty(100)[::2])
    with pytest.raises(ValueError):
        h.predict(xte, mean_out=np.empty(49))
    with pytest.raises(ValueError):
        h.kernel(xte, out=np.empty((50, 50), dtype=np.float32))
    m, v = np.empty(50), np.empty(50)
    h.predict(xte, mean_out=m, var_out=v)
    m2, v2 = h.predict(xte)
    assert np.array_equal(m, m2) and np.array_equal(v, v2)
    # float32 / strided INPUTS are still converted
    m3, _ = h.predict(np.asfortranarray(xte))
    assert np.array_equal(m3, m2)


def test_model_file_name_without_suffix(lib, synth, tmp_path):
    xtr, ytr, xte, _ = synth.make_problem(150, 40, 8)
    h = lib.Handle()
    h.fit(xtr, ytr)
    h.save(tmp_path / "model")                      # np.savez appends .npz
    h2 = lib.Handle.load(tmp_path / "model")
    assert np.array_equal(h2.predict(xte)[0], h.predict(xte)[0])


# ---------------------------------------------------------------------------------------------- multi-GPU
def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _need_two_gpus():
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")


def test_in_process_multi_gpu_is_bitwise_the_single_gpu_result(lib, synth):
    """One handle, G GPUs behind the C ABI: the fit is replicated peer-to-peer, nngp_predict splits the rows -- the
    result must be the 1-GPU result bit for bit, with host and with device buffers, and the
    replicas must follow every refit (append_fit, set_state)."""
    import torch
    _need_two_gpus()
    g = min(_n_gpus(), 8)
    xtr, ytr, xte, _ = synth.make_problem(3000, 5000, 40)
    h1 = lib.Handle(depth=3)
    h1.fit(xtr, ytr)
    m1, v1 = h1.predict(xte)
    hg = lib.Handle(depth=3, n_gpus=g)
    assert hg.n_gpus == g
    hg.fit(xtr, ytr)
    s = hg.stats()
    assert s["replicate_bytes"] >= 8 * 3000 * 3001 // 2 and s["replicate_ms"] > 0
    mg, vg = hg.predict(xte)
    assert np.array_equal(mg, m1) and np.array_equal(vg, v1)
    # odd split sizes, device-resident inputs / outputs on the first GPU
    xd = torch.from_numpy(xte[:4321]).cuda(0)
    md = torch.empty(4321, dtype=torch.float64, device="cuda:0")
    vd = torch.empty(4321, dtype=torch.float64, device="cuda:0")
    hg.predict(xd, mean_out=md, var_out=vd)
    assert np.array_equal(md.cpu().numpy(), m1[:4321]) and np.array_equal(vd.cpu().numpy(), v1[:4321])
    # a small batch stays on the first GPU
    ms, vs = hg.predict(xte[:5])
    assert np.array_equal(ms, m1[:5]) and np.array_equal(vs, v1[:5])
    # selection runs the multi-GPU prediction underneath
    assert np.array_equal(hg.active_select(xte, 64), h1.active_select(xte, 64))
    # refits propagate: append, then import into a fresh multi-GPU handle
    xn, yn = synth.encodings(500, 40, 9), synth.labels(synth.encodings(500, 40, 9))
    h1.append_fit(xn, yn)
    hg.append_fit(xn, yn)
    m1b, v1b = h1.predict(xte)
    mgb, vgb = hg.predict(xte)
    assert not np.array_equal(m1b, m1)
    assert np.array_equal(mgb, m1b) and np.array_equal(vgb, v1b)
    st = h1.get_state()
    hi = lib.Handle(depth=3, device_ids=list(range(g))[::-1])       # fit GPU = the last ordinal
    hi.set_state(st["x"], st["l"], st["alpha"], st["lambda"])
    mi, vi = hi.predict(xte)
    assert np.array_equal(mi, m1b) and np.array_equal(vi, v1b)
    with pytest.raises(ValueError):
        lib.Handle(device_ids=[0, 0])


def test_in_process_multi_gpu_ntk(lib, synth):
    _need_two_gpus()
    xtr, ytr, xte, _ = synth.make_problem(900, 700, 24)
    h1 = lib.Handle(kernel_type="ntk")
    h1.fit(xtr, ytr)
    m1, v1 = h1.predict(xte)
    hg = lib.Handle(kernel_type="ntk", n_gpus=2)
    hg.fit(xtr, ytr)
    mg, vg = hg.predict(xte)
    assert np.array_equal(mg, m1) and np.array_equal(vg, v1)


def test_nccl_broadcast_fit_gives_every_rank_the_same_bits():
    """the one-process-per-GPU path -- rank 0 fits, broadcast_fit ships the packed state over NCCL,
    every rank predicts the SAME rows; ranks > 0 must reproduce rank 0 bit for bit (a race between torch's
    communication stream and the library's stream would show up here)."""
    _need_two_gpus()
    g = min(_n_gpus(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={g}",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        str(ROOT / "tests" / "checks" / "nccl_broadcast_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "NCCL_BROADCAST_OK" in r.stdout


def test_symmetric_kernel_skips_upper_tiles_and_stays_bitwise_symmetric(lib, synth):
    """kernel_fn(x, None) (train.py:216): only the tiles at / below the diagonal are computed, the rest is mirrored --
    the result must be what the full rectangular computation kernel_fn(x, x) gives, bit for bit."""
    x = synth.encodings(1500, 40, 4)
    for depth in (2, 3):
        h = lib.Handle(depth=depth, stats_level=2)
        k_sym = h.kernel(x)
        flops_sym = h.stats()["gram_flops"]
        h.stats_reset()
        k_full = h.kernel(x, x)
        assert np.array_equal(k_sym, k_full) and np.array_equal(k_sym, k_sym.T)
        assert flops_sym < 0.6 * h.stats()["gram_flops"]
        assert relmax(k_sym, oracle.kernel_fn(x, None, depth)) < 1e-13


# ---------------------------------------------------------------------------------------------- per-layer sigmas
@pytest.mark.parametrize("case", ["layers_d2", "layers_d3"])
def test_per_layer_sigmas_match_mpmath_and_oracle(lib, golden_dir, case):
    """stax.serial chains whose Dense layers differ in W_std / b_std (nngp_config.per_layer): kernel entries against
    50-digit mpmath, posterior against mpmath and the oracle, NNGP and NTK; through the stax mirror too."""
    from nngp_b200 import stax
    z = np.load(golden_dir / "mp_layers.npz")
    g = {k.split("/")[1]: z[k] for k in z.files if k.startswith(case + "/")}
    sw, sb, reg = tuple(float(v) for v in g["sigma_w"]), tuple(float(v) for v in g["sigma_b"]), float(g["diag_reg"])
    depth = len(sw)
    x, y, xt = g["x_train"], g["y_train"], g["x_test"]
    h = lib.Handle(depth=depth, sigma_w=sw, sigma_b=sb, diag_reg=reg)
    assert relmax(h.kernel(x), g["K_dd"]) < 1e-13 and relmax(h.kernel(xt, x), g["K_td"]) < 1e-13
    h.fit(x, y)
    m, v = h.predict(xt)
    assert abs(h.dims()[2] - float(g["lam"])) < 1e-12 * float(g["lam"])
    assert relmax(m, g["mean"]) < 1e-8 and relmax(v, g["var"]) < 1e-7
    ref = oracle.Fit(x, y, depth, sw, sb, diag_reg=reg)
    rm, rv = ref.predict(xt)
    assert relmax(m, rm) < 1e-9 and relmax(v, rv) < 1e-8
    hn = lib.Handle(depth=depth, sigma_w=sw, sigma_b=sb, diag_reg=reg, kernel_type="ntk")
    assert relmax(hn.kernel(xt, x), g["Theta_td"]) < 1e-8
    hn.fit(x, y)
    mn, vn = hn.predict(xt)
    assert relmax(mn, g["ntk_mean"]) < 1e-6 and relmax(vn, g["ntk_var"]) < 1e-5
    # the stax mirror builds the same spec from the layer list
    layers = []
    for l in range(depth):
        layers += ([stax.Relu()] if l else []) + [stax.Dense(512 if l + 1 < depth else 1, W_std=sw[l], b_std=sb[l])]
    _, _, kernel_fn = stax.serial(*layers)
    assert np.array_equal(kernel_fn(x), h.kernel(x))
    # a model file keeps the per-layer values
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        h.save(td + "/m")
        h2 = lib.Handle.load(td + "/m")
        assert np.array_equal(h2.predict(xt)[0], m)
    with pytest.raises(ValueError):
        lib.Handle(depth=3, sigma_w=(1.0, 2.0))


def test_new_entry_points_reject_misuse(lib, synth):
    xtr, ytr, xte, _ = synth.make_problem(100, 10, 8)
    h = lib.Handle()
    with pytest.raises(lib.NngpError):
        h.state_pack(0, 4, np.empty(4))                       # nothing fitted
    with pytest.raises(lib.NngpError):
        h.state_import_end(1.0)                               # no import in progress
    with pytest.raises(ValueError):
        lib.Handle(n_gpus=9)
    with pytest.raises((lib.NngpError, ValueError)):
        lib.Handle(device_ids=[0, 99])                        # no such GPU
    h.fit(xtr, ytr)
    with pytest.raises(ValueError):
        h.state_pack(0, 10 ** 12, np.empty(4))
    assert h.n_gpus == 1 and len(lib.build_id()) == 16


def test_in_process_multi_gpu_latency_mode_and_reserve(lib, synth):
    """Replicas carry the explicit inverse factor (latency mode) and follow appends after nngp_reserve sized them."""
    _need_two_gpus()
    xtr, ytr, xte, _ = synth.make_problem(1500, 1200, 24)
    h1 = lib.Handle(latency_mode=True)
    hg = lib.Handle(latency_mode=True, n_gpus=2)
    hg.reserve(2500, 24, 1200)
    h1.fit(xtr, ytr)
    hg.fit(xtr, ytr)
    m1, v1 = h1.predict(xte)             # 1200 rows <= 4096: the inverse path, on one / on two GPUs
    mg, vg = hg.predict(xte)
    assert np.array_equal(mg, m1) and np.array_equal(vg, v1)
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xte[:200])
    assert relmax(mg[:200], rm) < 1e-6 and np.max(np.abs(vg[:200] - rv) / np.abs(rv)) < 1e-6
    xn = synth.encodings(700, 24, 9)
    yn = synth.labels(xn)
    h1.append_fit(xn, yn)
    hg.append_fit(xn, yn)
    m1b, v1b = h1.predict(xte)
    mgb, vgb = hg.predict(xte)
    assert np.array_equal(mgb, m1b) and np.array_equal(vgb, v1b) and not np.array_equal(v1b, v1)

"""GPU parity tests proper: the sm_100a CUDA path, called through the C ABI (ctypes), against the CPU
oracle on identical seeded inputs, the committed golden fixtures, and size-independent properties.

Tolerances (north_star): kernel entries and predicted log-cardinalities within 1e-6 relative in FP64
(1e-3 on q-error).  The asserted bounds below are tighter where FP64 re-association allows it.
"""
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


# ---------------------------------------------------------------------------------------------- kernel_fn
@pytest.mark.parametrize("m,n,d", [(1, 1, 1), (5, 7, 3), (64, 64, 16), (129, 65, 17), (300, 200, 20), (1000, 900, 130)])
def test_gram_only_depth1_is_exact_gemm(lib, m, n, d):
    rng = np.random.default_rng(m * 1000 + n)
    a, b = rng.standard_normal((m, d)), rng.standard_normal((n, d))
    h = lib.Handle(depth=1)
    assert relmax(h.kernel(a, b), a @ b.T / d) < 1e-13


@pytest.mark.parametrize("depth,sw,sb", [(2, 1.0, 0.0), (3, 1.0, 0.0), (2, 1.5, 0.05), (4, 1.2, 0.1)])
def test_kernel_entries_match_oracle(lib, synth, depth, sw, sb):
    x1, x2 = synth.encodings(700, 40, 3), synth.encodings(450, 40, 4)
    x1[5] = x1[4]; x1[6] = 0.0; x2[0] = x1[4]; x2[1] = 2 * x1[4]; x2[2] = 0.0   # duplicates / zero rows
    h = lib.Handle(depth=depth, sigma_w=sw, sigma_b=sb)
    assert relmax(h.kernel(x1, x2), oracle.kernel_fn(x1, x2, depth, sw, sb)) < 1e-12
    k = h.kernel(x1)
    assert relmax(k, oracle.kernel_fn(x1, None, depth, sw, sb)) < 1e-12
    assert np.array_equal(k, k.T) or relmax(k, k.T) < 1e-15    # symmetric input -> symmetric kernel


def test_kernel_analytic_known_answers(lib):
    h = lib.Handle(depth=2)
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1000, (6, 8))
    k = h.kernel(x)
    assert relmax(np.diag(k), np.einsum("ij,ij->i", x, x) / 16) < 1e-14          # K(x,x) = |x|^2/(2D)
    a = np.zeros((1, 8)); a[0, :4] = [1, 2, 3, 4]
    b = np.zeros((1, 8)); b[0, 4:] = [4, 3, 2, 1]
    assert relmax(h.kernel(a, b), np.sqrt((a @ a.T / 8) * (b @ b.T / 8)) / (2 * np.pi)) < 1e-15  # orthogonal
    assert abs(h.kernel(a, -a)[0, 0]) < 1e-15                                    # antiparallel -> 0
    assert relmax(h.kernel(a, 3.5 * a), 3.5 * (a @ a.T / 8) / 2) < 1e-15         # parallel
    z = np.zeros((2, 8))
    assert np.all(h.kernel(z, x) == 0.0)                                         # theta = pi/2 branch


@pytest.mark.parametrize("case", ["ref_d2", "d3", "sigma_variant", "d1_linear", "abs_reg", "degenerate"])
def test_against_mpmath_golden(lib, mp_golden, case):
    g = mp_golden[case]
    depth, sw, sb, reg, absolute = g["cfg"]
    h = lib.Handle(depth=int(depth), sigma_w=sw, sigma_b=sb, diag_reg=reg, diag_reg_absolute=bool(absolute))
    scale = np.max(np.abs(g["K_dd"]))
    assert np.max(np.abs(h.kernel(g["x_train"]) - g["K_dd"])) < 1e-13 * scale
    assert np.max(np.abs(h.kernel(g["x_test"], g["x_train"]) - g["K_td"])) < 1e-13 * scale
    h.fit(g["x_train"], g["y_train"])
    assert abs(h.dims()[2] - g["lam"]) < 1e-13 * abs(g["lam"])
    mean, var = h.predict(g["x_test"])
    assert np.max(np.abs(mean - g["mean"])) < 1e-8 * np.max(np.abs(g["mean"]))
    assert np.max(np.abs(var - g["var"])) < 1e-8 * np.max(np.abs(g["var"]))
    assert relmax(h.get_state()["alpha"], g["alpha"]) < 1e-7


# ---------------------------------------------------------------------------------------------- Cholesky
@pytest.mark.parametrize("n", [1, 8, 64, 65, 128, 200, 257, 513, 1000])
def test_blocked_cholesky_matches_lapack(lib, n):
    import scipy.linalg as sla
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n + 8))
    spd = a @ a.T + n * np.eye(n)
    l = lib.Handle().potrf(spd)
    assert relmax(l, sla.cholesky(spd, lower=True)) < 1e-12
    assert relmax(l @ l.T, spd) < 1e-13


def test_not_positive_definite_is_reported(lib):
    h = lib.Handle()
    with pytest.raises(np.linalg.LinAlgError):
        h.potrf(-np.eye(70))
    x = np.ones((40, 6))
    hh = lib.Handle(diag_reg=0.0)        # rank-1 kernel, no regulariser -> singular
    with pytest.raises(np.linalg.LinAlgError):
        hh.fit(x, np.ones(40))


# ---------------------------------------------------------------------------------------------- fit + predict
@pytest.mark.parametrize("n,t,d,depth", [(40, 12, 20, 2), (700, 300, 24, 2), (1500, 1000, 64, 3), (2500, 700, 128, 2),
                                         (130, 1, 7, 2)])
def test_fit_predict_match_oracle(lib, synth, n, t, d, depth):
    d2 = d + (d % 2)
    xtr, ytr, xte, _ = synth.make_problem(n, t, d2)
    xtr, xte = xtr[:, :d], xte[:, :d]                      # odd D exercises the padded operand path
    h = lib.Handle(depth=depth)
    h.fit(xtr, ytr)
    ref = oracle.Fit(xtr, ytr, depth)
    st = h.get_state()
    assert abs(st["lambda"] - ref.lam) < 1e-13 * ref.lam
    assert relmax(st["l"], ref.c) < 1e-9
    assert np.all(np.triu(st["l"], 1) == 0.0)
    mean, var = h.predict(xte)
    rm, rv = ref.predict(xte)
    assert relmax(mean, rm) < 1e-6 and relmax(var, rv) < 1e-6
    assert np.max(np.abs(2.0 ** np.abs(mean - rm) - 1.0)) < 1e-3          # q-error gate
    m_only, none = h.predict(xte, want_var=False)
    assert none is None and np.array_equal(m_only, mean)                   # compute_cov=False path


def test_forest_workload_parity(lib, forest):
    """Config C1: the reference's shipped forest queries, N=10800 / T=3600 / D=20 / depth 2."""
    h = lib.Handle()
    h.fit(forest["x_train"], forest["y_train"])
    ref = oracle.Fit(forest["x_train"], forest["y_train"])
    assert abs(h.dims()[2] - ref.lam) < 1e-12 * ref.lam
    mean, var = h.predict(forest["x_test"])
    rm, rv = ref.predict(forest["x_test"])
    assert relmax(mean, rm) < 1e-6
    assert np.max(np.abs(var - rv) / np.abs(rv)) < 1e-6
    assert np.max(np.abs(2.0 ** np.abs(mean - rm) - 1.0)) < 1e-3
    q, rq = oracle.q_error_stats(mean, forest["y_test"]), oracle.q_error_stats(rm, forest["y_test"])
    assert abs(q["median"] - rq["median"]) < 1e-3 * rq["median"]


def test_row_blocking_and_sharding_are_bitwise_invariant(lib, synth):
    """k-GPU == 1-GPU: a test row's result must not depend on which block / shard it is in."""
    xtr, ytr, xte, _ = synth.make_problem(900, 1000, 32)
    h = lib.Handle()
    h.fit(xtr, ytr)
    mean, var = h.predict(xte)
    small = lib.Handle(max_block_bytes=256 * 912 * 8)       # forces 256-row blocks
    st = h.get_state()
    small.set_state(st["x"], st["l"], st["alpha"], st["lambda"])
    m2, v2 = small.predict(xte)
    assert np.array_equal(mean, m2) and np.array_equal(var, v2)
    parts = [small.predict(xte[a:b]) for a, b in ((0, 123), (123, 700), (700, 1000))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), mean)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), var)


def test_small_batch_and_persistent_solve_paths_agree_bitwise(lib, synth, monkeypatch):
    """The right-looking small-batch path and the persistent fused kernel share their arithmetic: same bits."""
    xtr, ytr, xte, _ = synth.make_problem(1100, 700, 20)
    h = lib.Handle()
    h.fit(xtr, ytr)
    monkeypatch.setenv("NNGP_SMALL_BATCH_TILES", "0")          # force the persistent fused kernel
    m_f, v_f = h.predict(xte)
    monkeypatch.setenv("NNGP_SMALL_BATCH_TILES", "1000000")    # force the right-looking steps
    m_s, v_s = h.predict(xte)
    assert np.array_equal(m_f, m_s) and np.array_equal(v_f, v_s)
    one_f = h.predict(xte[:1])                                  # single query (the PostgreSQL serving case)
    assert one_f[0][0] == m_f[0] and one_f[1][0] == v_f[0]
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xte)
    assert relmax(m_s, rm) < 1e-6 and relmax(v_s, rv) < 1e-6


def test_pipelined_and_ungated_kernel_variants_agree_bitwise(lib, synth):
    """trsm_fused_kernel<PIPE=false> (many row tiles: one up-front wait per item) and <PIPE=true> (few row tiles:
    every operand tile gated on its producer, several CTAs pipelining along J inside a row tile) are the same
    arithmetic: a row's result does not depend on which variant -- i.e. on how many rows came with it."""
    xtr, ytr, _, _ = synth.make_problem(1100, 4, 20)
    xte = synth.encodings(48000, 20, 7)
    h = lib.Handle()
    h.fit(xtr, ytr)
    m_big, v_big = h.predict(xte)                      # 375 row tiles -> ungated variant
    for rows in (1, 130, 5000, 37888):                 # 1 .. 296 row tiles -> pipelined variant (296: static, ungated)
        m, v = h.predict(xte[:rows])
        assert np.array_equal(m, m_big[:rows]) and np.array_equal(v, v_big[:rows]), rows
    m_tail, v_tail = h.predict(xte[40000:])            # other rows, pipelined
    assert np.array_equal(m_tail, m_big[40000:]) and np.array_equal(v_tail, v_big[40000:])


def test_repeated_fits_are_bit_identical(lib, synth):
    """Race detector of last resort (compute-sanitizer is closed on this pool): the same fit + predict, repeated on
    three alternating handles, must give identical bits (this is what exposed the look-ahead corruption)."""
    xtr, ytr, xte, _ = synth.make_problem(4096, 1024, 64)
    handles = [lib.Handle() for _ in range(3)]
    ref = None
    for rep in range(9):
        h = handles[rep % 3]
        h.fit(xtr, np.ones_like(ytr))
        cur = (h.get_state(x=False, l=False)["alpha"],) + tuple(h.predict(xte))
        if ref is None:
            ref = cur
        else:
            assert all(np.array_equal(a, b) for a, b in zip(cur, ref)), f"fit #{rep} differs from fit #0"


def test_lookahead_schedule_is_bitwise_neutral(lib, synth, monkeypatch):
    """The two-stream look-ahead Cholesky (default) must give the factor of the single-stream schedule, every time
    (it overlaps the one-CTA panel kernels with DMMA GEMM CTAs on the same SMs -- the case that exposed the
    stage-release hazard)."""
    xtr, ytr, xte, _ = synth.make_problem(8192, 256, 64)
    monkeypatch.setenv("NNGP_CHOL_LOOKAHEAD", "0")
    h0 = lib.Handle()
    h0.fit(xtr, ytr)
    ref = h0.get_state(x=False)
    monkeypatch.delenv("NNGP_CHOL_LOOKAHEAD")
    h1 = lib.Handle()
    for rep in range(8):
        h1.fit(xtr, ytr)
        cur = h1.get_state(x=False)
        assert np.array_equal(cur["l"], ref["l"]), f"look-ahead fit #{rep}: factor differs from the single-stream one"
        assert np.array_equal(cur["alpha"], ref["alpha"])


def test_two_threads_match_the_serial_result(lib, synth):
    """Two handles driven from two host threads, their kernels genuinely overlapping on the GPU (no library-level
    lock): results must equal the serial reference bit for bit.  Regression test for the stage-release hazard of
    DESIGN.md section 5.3 (a co-resident CTA delayed a fragment LDS past the TMA refill of its ring stage)."""
    import threading
    xtr, ytr, xte, _ = synth.make_problem(4096, 2048, 64)
    y1 = np.ones_like(ytr)
    ref_h = lib.Handle()
    ref_h.fit(xtr, y1)
    ref = (ref_h.get_state(x=False, l=False)["alpha"],) + tuple(ref_h.predict(xte))
    errors = []

    def worker():
        try:
            h = lib.Handle()
            for _ in range(4):
                h.fit(xtr, y1)
                cur = (h.get_state(x=False, l=False)["alpha"],) + tuple(h.predict(xte))
                if not all(np.array_equal(a, b) for a, b in zip(cur, ref)):
                    errors.append("mismatch")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=worker) for _ in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_full_size_properties_c2(lib, synth):
    """BASELINE config C2 sizes (N=8192, D=128, depth 2): size-independent properties + sampled oracle check."""
    xtr, ytr, xte, _ = synth.make_problem(8192, 8192, 128)
    h = lib.Handle()
    h.fit(xtr, ytr)
    n, d, lam = h.dims()
    assert (n, d) == (8192, 128)
    assert abs(lam - 1e-3 * np.mean(np.einsum("ij,ij->i", xtr, xtr)) / 256) < 1e-12 * lam   # lambda from the diagonal
    # (a) predicting the training rows: K_* = K_dd  =>  mean = y - lambda*alpha, var = lambda - lambda^2 [A^-1]_ii
    alpha = h.get_state(x=False, l=False)["alpha"]
    mean_tr, var_tr = h.predict(xtr)
    assert np.max(np.abs(mean_tr - (ytr - lam * alpha))) < 1e-6 * np.max(np.abs(ytr))
    assert np.all(var_tr > 0) and np.all(var_tr < lam * (1 + 1e-9))
    # (b) linearity of the posterior mean in y
    h2 = lib.Handle()
    h2.fit(xtr, 3.0 * ytr + 1.0)
    h3 = lib.Handle()
    h3.fit(xtr, np.ones_like(ytr))
    m1, v1 = h.predict(xte[:2048])
    m2, v2 = h2.predict(xte[:2048])
    m3, _ = h3.predict(xte[:2048])
    assert np.max(np.abs(m2 - (3.0 * m1 + m3))) < 1e-6 * np.max(np.abs(m2))
    assert np.array_equal(v1, v2)                                            # variance does not depend on y
    # (c) sampled oracle parity at full N
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xte[:512])
    assert relmax(m1[:512], rm) < 1e-6 and np.max(np.abs(v1[:512] - rv) / np.abs(rv)) < 1e-6


def test_full_size_properties_depth3_16k(lib, synth):
    """C5-like sizes (N=16384, D=512 with 96 join dims, depth 3; W=512 Cholesky panels): the oracle needs minutes
    here, so parity rests on size-independent identities -- lambda from the computed diagonal, the posterior at the
    training rows (mean = y - lambda*alpha, 0 < var < lambda), the closed-form diagonal q_L = q_0 / 2^(L-1), and
    the persistent kernel against the right-looking path on 40 000 rows (313 row tiles), bit for bit."""
    xtr, _, xte, _ = synth.make_problem(16384, 40000, 512, join_dims=96)
    ytr = np.random.default_rng(3).uniform(0.0, 20.0, 16384)      # (the synthetic label saturates at this width)
    h = lib.Handle(depth=3)
    h.fit(xtr, ytr)
    n, d, lam = h.dims()
    q0 = np.einsum("ij,ij->i", xtr, xtr) / 512
    assert (n, d) == (16384, 512) and abs(lam - 1e-3 * np.mean(q0) / 4) < 1e-12 * lam
    alpha = h.get_state(x=False, l=False)["alpha"]
    mean_tr, var_tr = h.predict(xtr)
    assert np.max(np.abs(mean_tr - (ytr - lam * alpha))) < 1e-6 * np.max(np.abs(ytr))
    assert np.all(var_tr > 0) and np.all(var_tr < lam * (1 + 1e-9))
    kd = np.diag(h.kernel(xtr[:300]))
    assert np.max(np.abs(kd - q0[:300] / 4)) < 1e-12 * np.max(kd)
    import os
    m_f, v_f = h.predict(xte)                                     # 313 row tiles: persistent fused kernel
    os.environ["NNGP_SMALL_BATCH_TILES"] = "1000000"
    try:
        m_s, v_s = h.predict(xte[:6000])                          # right-looking path on a slice
    finally:
        del os.environ["NNGP_SMALL_BATCH_TILES"]
    assert np.array_equal(m_f[:6000], m_s) and np.array_equal(v_f[:6000], v_s)
    assert np.all(np.isfinite(m_f)) and np.all(v_f > 0)


def test_state_roundtrip_and_device_pointers(lib, synth):
    import torch
    xtr, ytr, xte, _ = synth.make_problem(600, 300, 16)
    h = lib.Handle()
    h.fit(xtr, ytr)
    mean, var = h.predict(xte)
    st = h.get_state()
    h2 = lib.Handle()
    h2.set_state(st["x"], st["l"], st["alpha"], st["lambda"])
    m2, v2 = h2.predict(xte)
    assert np.array_equal(mean, m2) and np.array_equal(var, v2)
    # device-resident inputs/outputs (what the sharded predictor and bench.py use)
    xd = torch.from_numpy(xte).cuda()
    md, vd = torch.empty(300, dtype=torch.float64, device="cuda"), torch.empty(300, dtype=torch.float64, device="cuda")
    h.predict(xd, mean_out=md, var_out=vd)
    assert np.array_equal(md.cpu().numpy(), mean) and np.array_equal(vd.cpu().numpy(), var)


# ---------------------------------------------------------------------------------------------- NTK mode (row f-2)
@pytest.mark.parametrize("depth,sw,sb", [(2, 1.0, 0.0), (3, 1.5, 0.05)])
def test_ntk_kernel_entries(lib, synth, depth, sw, sb):
    x1, x2 = synth.encodings(400, 24, 3), synth.encodings(300, 24, 4)
    h = lib.Handle(depth=depth, sigma_w=sw, sigma_b=sb, kernel_type="ntk")
    # no duplicate rows here: Theta is first-order sensitive to theta at theta = 0 (see tests/test_oracle.py)
    assert relmax(h.kernel(x1, x2), oracle.kernel_fn(x1, x2, depth, sw, sb, get="ntk")) < 1e-12
    th = h.kernel(x1)
    ref = oracle.kernel_fn(x1, None, depth, sw, sb, get="ntk")
    off = ~np.eye(400, dtype=bool)
    assert np.max(np.abs(th - ref)[off]) < 1e-12 * np.max(ref)
    assert np.max(np.abs(np.diag(th) - np.diag(ref))) < 1e-7 * np.max(ref)


@pytest.mark.parametrize("case", ["ntk_d2", "ntk_d3_sigma"])
def test_ntk_against_mpmath_golden(lib, mp_golden, case):
    g = mp_golden[case]
    depth, sw, sb, reg, _ = g["cfg"]
    h = lib.Handle(depth=int(depth), sigma_w=sw, sigma_b=sb, diag_reg=reg, kernel_type="ntk")
    h.fit(g["x_train"], g["y_train"])
    assert abs(h.dims()[2] - g["lam"]) < 1e-8 * abs(g["lam"])
    mean, var = h.predict(g["x_test"])
    assert np.max(np.abs(mean - g["mean"])) < 1e-7 * np.max(np.abs(g["mean"]))
    assert np.max(np.abs(var - g["var"])) < 1e-7 * np.max(np.abs(g["var"]))


@pytest.mark.parametrize("n,t,d,depth", [(600, 300, 24, 2), (1500, 700, 64, 3)])
def test_ntk_fit_predict_match_oracle(lib, synth, n, t, d, depth):
    """train.py --kernel_type ntk: predict_fn(get='ntk', compute_cov=True) -> mean and diag(cov) (Appendix A.5)."""
    xtr, ytr, xte, _ = synth.make_problem(n, t, d)
    h = lib.Handle(depth=depth, kernel_type="ntk")
    h.fit(xtr, ytr)
    ref = oracle.FitNTK(xtr, ytr, depth)
    assert abs(h.dims()[2] - ref.lam) < 1e-8 * ref.lam
    mean, var = h.predict(xte)
    rm, rv = ref.predict(xte)
    assert relmax(mean, rm) < 1e-6 and relmax(var, rv) < 1e-6
    m_only, _ = h.predict(xte, want_var=False)
    assert np.array_equal(m_only, mean)
    small = lib.Handle(depth=depth, kernel_type="ntk", max_block_bytes=2 * 128 * ((n + 15) // 16 * 16) * 8)
    small.fit(xtr, ytr)
    m2, v2 = small.predict(xte)                      # 128-row blocks: bitwise invariant to the blocking
    assert np.array_equal(m2, mean) and np.array_equal(v2, var)
    assert set(h.get_state(x=False)) == {"l", "alpha", "m", "lambda"}     # 'ntk' state also carries M


def test_ntk_state_roundtrip_and_model_file(lib, synth, tmp_path):
    """'ntk' state = {X, L, alpha, lambda, M}: export -> import on a fresh handle (what broadcast_fit does on the other
    ranks) and save -> load give bitwise the same predictions; without M an imported state serves the mean only."""
    xtr, ytr, xte, _ = synth.make_problem(500, 200, 16)
    h = lib.Handle(kernel_type="ntk")
    h.fit(xtr, ytr)
    mean, var = h.predict(xte)
    st = h.get_state()
    assert set(st) == {"x", "l", "alpha", "m", "lambda"} and st["m"].shape == (500, 500)
    assert np.allclose(st["m"], st["m"].T, rtol=0, atol=1e-9 * np.max(np.abs(st["m"])))
    h2 = lib.Handle(kernel_type="ntk")
    h2.set_state(st["x"], st["l"], st["alpha"], st["lambda"])
    m_only, _ = h2.predict(xte, want_var=False)
    assert np.array_equal(m_only, mean)
    with pytest.raises(lib.NngpError):
        h2.predict(xte)                                     # variance needs M
    h2.set_state(st["x"], st["l"], st["alpha"], st["lambda"], m=st["m"])
    m2, v2 = h2.predict(xte)
    assert np.array_equal(m2, mean) and np.array_equal(v2, var)
    h.save(tmp_path / "ntk.npz")
    h3 = lib.Handle.load(tmp_path / "ntk.npz")
    assert h3.is_ntk
    m3, v3 = h3.predict(xte)
    assert np.array_equal(m3, mean) and np.array_equal(v3, var)


def test_log_marginal_likelihood_and_model_file(lib, synth, tmp_path):
    xtr, ytr, xte, _ = synth.make_problem(900, 200, 24)
    best = None
    for depth, sw, sb in [(2, 1.0, 0.0), (3, 1.0, 0.0), (2, 1.5, 0.05)]:
        h = lib.Handle(depth=depth, sigma_w=sw, sigma_b=sb)
        h.fit(xtr, ytr)
        ref = oracle.Fit(xtr, ytr, depth, sw, sb)
        lml, rl = h.log_marginal_likelihood(), ref.log_marginal_likelihood()
        assert abs(lml - rl) < 1e-9 * abs(rl)
        best = max(best or (lml, depth), (lml, depth))
    assert best is not None
    h.save(tmp_path / "model.npz")
    mean, var = h.predict(xte)
    h2 = lib.Handle.load(tmp_path / "model.npz")
    m2, v2 = h2.predict(xte)
    assert np.array_equal(mean, m2) and np.array_equal(var, v2)
    with pytest.raises(lib.NngpError):
        h2.log_marginal_likelihood()          # an imported state carries no evidence terms


def test_active_learning_round_selects_the_oracle_rows(lib, synth):
    """Config C4 on the real engine: one round of active/ActiveLearner.py:43-77 (deterministic top-k branch)."""
    from nngp_b200 import stax
    from nngp_b200.active import ActiveLearner
    xtr, ytr, xpool, ypool = synth.make_problem(400, 900, 16)
    _, _, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    al = ActiveLearner(budget=100, active_iters=1, verbose=False)
    pf = al.train(kernel_fn, xtr, ytr[:, None])
    idx = al.active_test(pf, xpool)
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xpool)
    want = oracle.active_select(rm[:, None], np.sqrt(rv), 100)
    assert set(idx.tolist()) == set(want.tolist())
    pf2, x_end, y_end = al.active_train(kernel_fn, xtr, ytr[:, None], xpool, ypool[:, None], xpool[:50], ypool[:50, None])
    assert x_end.shape == (500, 16) and len(al.history) == 2


def test_device_selection_is_the_argsort_tail(lib, synth):
    """nngp_active_select == np.argsort(std / max(mean))[-budget:] (ActiveLearner.py:46-54): same rows
    as the oracle, returned in ascending score order, ties (duplicated pool rows) resolved like a stable sort, and
    budget >= T returns every row."""
    xtr, ytr, xpool, _ = synth.make_problem(1024, 5000, 32)
    xpool[100:164] = xpool[4000:4064]                     # 64 exact ties
    h = lib.Handle()
    h.fit(xtr, ytr)
    idx, score = h.active_select(xpool, 300, return_scores=True)
    assert idx.dtype == np.int64 and idx.shape == (300,)
    assert np.array_equal(idx, np.argsort(score, kind="stable")[-300:])          # exact radix select, exact order
    ref = oracle.Fit(xtr, ytr)
    rm, rv = ref.predict(xpool)
    ref_score = np.sqrt(rv) / np.max(rm)
    assert relmax(score, ref_score) < 1e-7
    want = oracle.active_select(rm[:, None], np.sqrt(rv), 300)
    # same rows as the oracle, up to members whose scores are equal within the parity tolerance of the scores
    diff = set(idx.tolist()) ^ set(want.tolist())
    cut = np.sort(ref_score)[-300]
    assert all(abs(ref_score[i] - cut) <= 1e-7 * cut for i in diff), sorted(diff)
    # a tie straddling nothing: the duplicated rows have identical device scores
    assert np.array_equal(score[100:164], score[4000:4064])
    # budget >= T: every row, still in argsort order; tiny pools work too
    idx_all = h.active_select(xpool[:77], 500)
    s77 = h.active_select(xpool[:77], 500, return_scores=True)[1]
    assert np.array_equal(idx_all, np.argsort(s77, kind="stable"))
    assert np.array_equal(h.active_select(xpool[:1], 3), [0])
    # the k = 1 and k = T - 1 ends of the radix select
    assert h.active_select(xpool, 1)[0] == np.argsort(score, kind="stable")[-1]
    assert np.array_equal(h.active_select(xpool, 4999), np.argsort(score, kind="stable")[-4999:])


def test_device_sampling_is_gumbel_top_k(lib, synth):
    """biased_sample=True branch (ActiveLearner.py:49-53) on the device: Gumbel-top-k over log(score) with the
    oracle's splitmix64 stream -> the same draw as the oracle's restatement (log() may differ in the last ulp, so a
    draw whose keys are within 1e-12 of the cut may swap), deterministic per seed, no repeats."""
    xtr, ytr, xpool, _ = synth.make_problem(512, 3000, 16)
    h = lib.Handle()
    h.fit(xtr, ytr)
    idx, score = h.active_select(xpool, 200, biased_sample=True, seed=10, return_scores=True)
    assert len(set(idx.tolist())) == 200 and idx.min() >= 0 and idx.max() < 3000
    assert np.array_equal(idx, h.active_select(xpool, 200, biased_sample=True, seed=10))
    assert not np.array_equal(idx, h.active_select(xpool, 200, biased_sample=True, seed=11))
    want = oracle.active_sample(np.ones(1), score, 200, seed=10)          # score is already std / max(mean)
    assert len(set(idx.tolist()) ^ set(want.tolist())) <= 2
    assert np.mean(idx == want) > 0.97                                    # draw order too
    # larger scores are drawn more often: mean score of the draw exceeds the pool's
    assert score[idx].mean() > score.mean()


def test_append_fit_is_bitwise_a_fresh_fit(lib, synth):
    """nngp_append_fit (merge_data + train, ActiveLearner.py:57-65,76) == nngp_fit on the stacked arrays, bit for bit;
    needs the labels of a previous nngp_fit on the same handle."""
    xtr, ytr, xpool, ypool = synth.make_problem(700, 300, 24)
    h = lib.Handle()
    h.reserve(900, 24, 300)                       # nngp_reserve: buffers sized once for the whole loop
    h.fit(xtr, ytr)
    h.append_fit(xpool[:150], ypool[:150])
    h.append_fit(xpool[150:151], ypool[150:151])
    ref = lib.Handle()
    ref.fit(np.vstack([xtr, xpool[:151]]), np.concatenate([ytr, ypool[:151]]))
    a, b = h.get_state(), ref.get_state()
    assert h.dims() == ref.dims() and h.dims()[0] == 851
    for k in ("x", "l", "alpha"):
        assert np.array_equal(a[k], b[k]), k
    assert h.log_marginal_likelihood() == ref.log_marginal_likelihood()
    with pytest.raises(ValueError):
        h.append_fit(xpool[:3, :5], ypool[:3])
    h2 = lib.Handle()
    st = ref.get_state()
    h2.set_state(st["x"], st["l"], st["alpha"], st["lambda"])
    with pytest.raises(lib.NngpError):                                    # imported state carries no labels
        h2.append_fit(xpool[:3], ypool[:3])
    h.reserve(2000, 24)                                                   # re-sizing drops the fitted model
    with pytest.raises(lib.NngpError):
        h.predict(xpool[:3])


def test_incremental_append_with_absolute_lambda(lib, synth, monkeypatch):
    """diag_reg_absolute_scale=True: nngp_append_fit extends the factor (block Cholesky of the Schur complement, N^2 M
    flop) instead of refactoring -- same model as a fresh fit up to rounding, and the oracle's; N is deliberately not
    a multiple of 16 / 64 and a one-row append is included."""
    xtr, ytr, xpool, ypool = synth.make_problem(900, 400, 24)
    lam = 50.0
    h = lib.Handle(diag_reg=lam, diag_reg_absolute=True)
    h.fit(xtr, ytr)
    h.append_fit(xpool[:130], ypool[:130])
    h.append_fit(xpool[130:131], ypool[130:131])
    xs, ys = np.vstack([xtr, xpool[:131]]), np.concatenate([ytr, ypool[:131]])
    fresh = lib.Handle(diag_reg=lam, diag_reg_absolute=True)
    fresh.fit(xs, ys)
    a, b = h.get_state(), fresh.get_state()
    assert h.dims() == fresh.dims() == (1031, 24, lam)
    assert np.array_equal(a["x"], b["x"])
    assert relmax(a["l"], b["l"]) < 1e-11 and relmax(a["alpha"], b["alpha"]) < 1e-8
    assert abs(h.log_marginal_likelihood() - fresh.log_marginal_likelihood()) < 1e-9 * abs(fresh.log_marginal_likelihood())
    m1, v1 = h.predict(xpool[200:])
    m2, v2 = fresh.predict(xpool[200:])
    assert relmax(m1, m2) < 1e-9 and relmax(v1, v2) < 1e-9
    ref = oracle.Fit(xs, ys, 2, 1.0, 0.0, lam, True)
    rm, rv = ref.predict(xpool[200:])
    assert relmax(m1, rm) < 1e-6 and relmax(v1, rv) < 1e-6
    # a third append works on the extended state (N = 1031 is odd: that one refits, the block needs 16-byte
    # alignment), and the switch forces the bitwise refit
    h.append_fit(xpool[131:195], ypool[131:195])
    monkeypatch.setenv("NNGP_APPEND_INCREMENTAL", "0")
    fresh.append_fit(xpool[131:195], ypool[131:195])
    full = lib.Handle(diag_reg=lam, diag_reg_absolute=True)
    full.fit(np.vstack([xs, xpool[131:195]]), np.concatenate([ys, ypool[131:195]]))
    assert np.array_equal(fresh.get_state()["l"], full.get_state()["l"])
    assert relmax(h.get_state()["alpha"], full.get_state()["alpha"]) < 1e-8


def test_active_learner_device_path_equals_host_path(lib, synth):
    """ActiveLearner.active_train through nngp_active_select / nngp_append_fit ends with exactly the training set the
    host-side restatement of the loop (numpy argsort + fresh fits) ends with."""
    from nngp_b200 import stax
    from nngp_b200.active import ActiveLearner
    xtr, ytr, xpool, ypool = synth.make_problem(300, 1200, 16)
    _, _, kernel_fn = stax.serial(stax.Dense(512), stax.Relu(), stax.Dense(1))
    al = ActiveLearner(budget=128, active_iters=3, verbose=False)
    pf, x_end, y_end = al.active_train(kernel_fn, xtr, ytr[:, None], xpool, ypool[:, None], xpool[:40], ypool[:40, None])
    host = ActiveLearner(budget=128, active_iters=3, verbose=False)
    xh, yh, xp, yp = xtr, ytr[:, None], xpool, ypool[:, None]
    for _ in range(3):
        pfh = host.train(kernel_fn, xh, yh)
        sel = host._active_test_host(pfh, xp)
        xh, yh, xp, yp = host.merge_data(sel, xh, yh, xp, yp)
    assert x_end.shape == (300 + 3 * 128, 16)
    assert np.array_equal(np.sort(x_end.view(np.uint64), axis=0), np.sort(xh.view(np.uint64), axis=0))
    # the returned predict_fn is bound to the appended engine: same predictions as a fresh fit on the final set
    fresh = host.train(kernel_fn, x_end, y_end)
    m1, c1 = pf(x_test=xpool[:64], get="nngp", compute_cov=True)
    m2, c2 = fresh(x_test=xpool[:64], get="nngp", compute_cov=True)
    assert np.array_equal(m1, m2) and np.array_equal(np.diag(c1), np.diag(c2))


def test_error_conventions(lib, synth):
    h = lib.Handle()
    xtr, ytr, xte, _ = synth.make_problem(64, 8, 8)
    with pytest.raises(lib.NngpError):
        h.predict(xte)                                   # not fitted
    bad = xtr.copy(); bad[3, 2] = np.nan
    with pytest.raises(ValueError):
        h.fit(bad, ytr)                                  # non-finite input
    with pytest.raises(ValueError):
        h.fit(xtr, ytr[:-1])                             # shape mismatch
    with pytest.raises(ValueError):
        lib.Handle(depth=0)
    h.fit(xtr, ytr)
    with pytest.raises(ValueError):
        h.predict(np.full((4, 8), np.inf))
    mean, var = h.predict(xte)                           # handle still usable after errors
    assert np.all(np.isfinite(mean))

"""CPU: pins the numpy/scipy oracle (the parity checker) -- no GPU, no CUDA library involved.

The reference ships no tests or golden vectors for this path, so the pins are:
analytic known answers, the 50-digit mpmath restatement (tests/golden/mp_small.npz), a finite-width
Monte-Carlo check of the /D and no-bias conventions, and the forest-workload invariants.
"""
import numpy as np
import pytest

import nngp_oracle as o


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def test_analytic_known_answers():
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1000, (6, 8))
    d = x.shape[1]
    k = o.kernel_fn(x, None, depth=2)
    # K(x,x) = |x|^2 / (2D)
    assert rel(np.diag(k), np.einsum("ij,ij->i", x, x) / (2 * d)) < 1e-14
    # orthogonal rows: K = sqrt(q q') / (2 pi)
    a = np.zeros((1, 8)); a[0, :4] = rng.uniform(1, 9, 4)
    b = np.zeros((1, 8)); b[0, 4:] = rng.uniform(1, 9, 4)
    qa, qb = a @ a.T / 8, b @ b.T / 8
    assert rel(o.kernel_fn(a, b), np.sqrt(qa * qb) / (2 * np.pi)) < 1e-15
    # x' = -x: K = 0 ; x' = c x: K = c |x|^2 / (2D)
    assert abs(o.kernel_fn(a, -a)[0, 0]) < 1e-15 * qa[0, 0]
    assert rel(o.kernel_fn(a, 3.5 * a), 3.5 * qa / 2) < 1e-15
    # general closed form: sqrt(q q') (sin t + (pi - t) cos t) / (2 pi)
    u, v = x[:1], x[1:2]
    qu, qv = (u @ u.T / d)[0, 0], (v @ v.T / d)[0, 0]
    t = np.arccos((u @ v.T / d)[0, 0] / np.sqrt(qu * qv))
    assert rel(o.kernel_fn(u, v)[0, 0], np.sqrt(qu * qv) * (np.sin(t) + (np.pi - t) * np.cos(t)) / (2 * np.pi)) < 1e-12
    # depth-L diagonal q_L = q_0 / 2^(L-1); zero rows hit the theta = pi/2 branch and give 0
    assert rel(np.diag(o.kernel_fn(x, None, depth=4)), np.einsum("ij,ij->i", x, x) / d / 8) < 1e-14
    z = np.zeros((2, 8))
    assert np.all(o.kernel_fn(z, x) == 0.0) and np.all(np.isfinite(o.kernel_fn(z, z)))


@pytest.mark.parametrize("case", ["ref_d2", "d3", "sigma_variant", "d1_linear", "abs_reg", "degenerate"])
def test_oracle_matches_mpmath_golden(mp_golden, case):
    g = mp_golden[case]
    depth, sw, sb, reg, absolute = g["cfg"]
    depth, absolute = int(depth), bool(absolute)
    kdd = o.kernel_fn(g["x_train"], None, depth, sw, sb)
    ktd = o.kernel_fn(g["x_test"], g["x_train"], depth, sw, sb)
    scale = np.max(np.abs(g["K_dd"]))
    # entries: 1e-14 relative to the kernel scale (near-duplicate rows lose relative accuracy in s, not in K)
    assert np.max(np.abs(kdd - g["K_dd"])) < 1e-14 * scale
    assert np.max(np.abs(ktd - g["K_td"])) < 1e-14 * scale
    fit = o.Fit(g["x_train"], g["y_train"], depth, sw, sb, reg, absolute)
    assert abs(fit.lam - g["lam"]) < 1e-14 * abs(g["lam"])
    mean, var = fit.predict(g["x_test"])
    assert np.max(np.abs(mean - g["mean"])) < 1e-9 * np.max(np.abs(g["mean"]))
    assert np.max(np.abs(var - g["var"])) < 1e-9 * np.max(np.abs(g["var"]))
    m2, cov = fit.predict_full_cov(g["x_test"])
    assert m2.shape == (len(mean), 1) and cov.shape == (len(mean), len(mean))
    assert np.max(np.abs(np.diag(cov) - g["var"])) < 1e-8 * np.max(np.abs(g["var"]))


@pytest.mark.parametrize("case", ["ntk_d2", "ntk_d3_sigma"])
def test_ntk_oracle_matches_mpmath_golden(mp_golden, case):
    g = mp_golden[case]
    depth, sw, sb, reg, _ = g["cfg"]
    depth = int(depth)
    scale = np.max(np.abs(g["Theta_dd"]))
    # Unlike K (dK/dtheta = 0 at theta = 0), Theta is first-order sensitive to theta through kdot = 1/2 - theta/(2 pi):
    # for duplicate rows / the diagonal, s = sqrt(q q' - k^2) is sqrt(rounding noise) ~ 1e-8 sqrt(qq') in ANY FP64
    # evaluation of the nt formula, so entries agree with the 50-digit value to ~1e-9 only there, 1e-14 elsewhere.
    dth = np.abs(o.kernel_fn(g["x_train"], None, depth, sw, sb, get="ntk") - g["Theta_dd"])
    assert np.max(dth) < 1e-8 * scale and np.median(dth) < 1e-14 * scale
    dts = np.abs(o.kernel_fn(g["x_test"], g["x_train"], depth, sw, sb, get="ntk") - g["Theta_td"])
    assert np.max(dts) < 1e-8 * scale and np.median(dts) < 1e-14 * scale
    fit = o.FitNTK(g["x_train"], g["y_train"], depth, sw, sb, reg)
    assert abs(fit.lam - g["lam"]) < 1e-9 * abs(g["lam"])
    mean, var = fit.predict(g["x_test"])
    assert np.max(np.abs(mean - g["mean"])) < 1e-8 * np.max(np.abs(g["mean"]))
    assert np.max(np.abs(var - g["var"])) < 1e-8 * np.max(np.abs(g["var"]))
    # Theta(x,x) closed form: kdot = 1/2 on the diagonal
    q0 = o.layer0_diag(g["x_train"], sw, sb)
    assert np.allclose(np.diag(g["Theta_dd"]), o.final_diag_ntk(q0, depth, sw, sb), rtol=1e-13)   # exact in mpmath


def test_finite_width_monte_carlo_conventions():
    """f(x) = v . relu(W x / sqrt(D)) / sqrt(width): covariance -> K; guards /D, W_std=1, no bias."""
    rng = np.random.default_rng(5)
    d, width, nets = 6, 2048, 200
    x = rng.uniform(0, 3, (4, d))
    acc = np.zeros((4, 4))
    for _ in range(nets):
        w = rng.standard_normal((width, d))
        v = rng.standard_normal(width)
        f = (np.maximum(x @ w.T / np.sqrt(d), 0.0) * v).sum(1) / np.sqrt(width)
        acc += np.outer(f, f)
    emp = acc / nets
    k = o.kernel_fn(x, None, depth=2)
    assert np.max(np.abs(emp - k)) / np.max(k) < 0.2  # 200 draws: ~10% sampling noise
    # a sharper check through the closed-form expectation of relu(u)relu(u') per hidden unit
    w = rng.standard_normal((400000, d))
    h = np.maximum(x @ w.T / np.sqrt(d), 0.0)
    assert np.max(np.abs(h @ h.T / w.shape[0] - k)) / np.max(k) < 1e-2


def test_forest_workload_invariants(forest):
    xtr, ytr, xte, yte = forest["x_train"], forest["y_train"], forest["x_test"], forest["y_test"]
    assert xtr.shape == (10800, 20) and xte.shape == (3600, 20) and xtr.dtype == np.float64
    # lambda = 1e-3 * trace(K)/N = 1e-3 * mean(|x|^2) / (2 D): independent of the solver
    lam = 1e-3 * np.mean(np.einsum("ij,ij->i", xtr, xtr) / 40.0)
    assert abs(lam - 185.5609) < 1e-3     # SURVEY.md section 6 probe value
    sub = slice(0, 1500)
    fit = o.Fit(xtr[sub], ytr[sub])
    mean, var = fit.predict(xte[:300])
    assert np.all(var > 0) and np.all(np.isfinite(mean))
    qe = o.q_error_stats(mean, yte[:300])
    assert qe["median"] < 10.0            # the estimator is doing something sensible


def test_active_selection_rule():
    mean = np.array([[4.0], [8.0], [2.0], [6.0]])
    std = np.array([0.4, 0.1, 0.9, 0.2])
    assert list(o.active_select(mean, std, 2)) == [0, 2]
    assert sorted(o.active_select(mean, std, 10)) == [0, 1, 2, 3]


def test_active_sampling_restatement():
    """Gumbel-top-k restatement of the biased_sample branch (ActiveLearner.py:49-53): splitmix64 known answer, no
    repeats, determinism, zero-probability rows never drawn, inclusion frequencies like numpy's choice(p=...)."""
    u = o.splitmix_uniform(0, 4)
    assert int(u[0] * 2.0 ** 52 - 0.5) == 0xE220A8397B1DCDAF >> 12      # first splitmix64 output for state 0
    assert np.all((u > 0) & (u < 1))
    rng = np.random.default_rng(0)
    s = rng.random(40) + 0.01
    s[7] = 0.0
    a = o.active_sample(np.ones(1), s, 5, seed=3)
    assert len(set(a.tolist())) == 5 and np.array_equal(a, o.active_sample(np.ones(1), s, 5, seed=3))
    cnt, cnt_np = np.zeros(40), np.zeros(40)
    p = s / s.sum()
    for seed in range(3000):
        cnt[o.active_sample(np.ones(1), s, 5, seed)] += 1
        cnt_np[np.random.default_rng(seed).choice(40, 5, replace=False, p=p)] += 1
    assert cnt[7] == 0
    assert np.max(np.abs(cnt - cnt_np)) / 3000 < 0.03 and np.corrcoef(cnt, cnt_np)[0, 1] > 0.98
    assert sorted(o.active_sample(np.ones(1), s[:4], 10, seed=1).tolist()) == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        o.active_sample(np.ones(1), -s, 3)


def test_device_atan2_algorithm_in_numpy():
    """The Gram epilogue's atan2_pos (csrc/gemm_nt.cuh) restated in numpy with the SAME coefficients (parsed from
    the source): <= 1 ulp-level agreement with libm and with 40-digit mpmath, incl. the axes and the origin."""
    import re
    from pathlib import Path
    import mpmath as mp
    src = (Path(__file__).resolve().parents[1] / "nngp-src_b200" / "csrc" / "gemm_nt.cuh").read_text()
    body = src[src.index("constexpr double C[21] = {"):]
    body = body[:body.index("};")]
    coef = [float(v) for v in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", body.split("{", 1)[1])]
    assert len(coef) == 21 and coef[0] == 1.0

    def atan2_pos(s, k):
        # min/max = s |k| / max(s^2, k^2) (reciprocal + product instead of a division; the device computes the
        # reciprocal with a MUFU seed + two Newton steps, <= 1 ulp from the correctly rounded one used here),
        # P(u) = Pe(u^2) + u Po(u^2): two Horner chains of 10
        a = np.abs(k)
        s2, a2 = s * s, k * k
        mx2 = np.maximum(s2, a2)
        with np.errstate(invalid="ignore", divide="ignore"):
            t = (s * a) * (1.0 / mx2)
        u = t * t
        w = u * u
        pe = np.full_like(u, coef[20])
        for c in coef[18::-2]:
            pe = pe * w + c
        po = np.full_like(u, coef[19])
        for c in coef[17::-2]:
            po = po * w + c
        at = t * (u * po + pe)
        th0 = np.where(s2 > a2, np.pi / 2 - at, at)
        th = np.where(k < 0, np.pi - th0, th0)
        return np.where(mx2 == 0, np.pi / 2, th)

    rng = np.random.default_rng(1)
    # magnitudes a kernel value can take (the squares must stay inside the FP64 range)
    s = np.abs(rng.standard_normal(200000)) * 10 ** rng.uniform(-12, 9, 200000)
    k = rng.standard_normal(200000) * 10 ** rng.uniform(-12, 9, 200000)
    s[:500] = 0.0; k[500:1000] = 0.0; s[1000] = k[1000] = 0.0; k[1001:1500] = s[1001:1500]
    ref = np.arctan2(s, k)
    ref[(s == 0) & (k == 0)] = np.pi / 2
    assert np.max(np.abs(atan2_pos(s, k) - ref)) <= 9e-16            # <= 2 ulp of theta in [0, pi]
    mp.mp.dps = 40
    hi = np.array([float(mp.atan2(mp.mpf(float(a)), mp.mpf(float(b)))) for a, b in zip(s[2000:4000], k[2000:4000])])
    assert np.max(np.abs(atan2_pos(s[2000:4000], k[2000:4000]) - hi)) <= 9e-16


@pytest.mark.parametrize("case", ["layers_d2", "layers_d3"])
def test_per_layer_sigmas_against_mpmath(case, golden_dir):
    """Dense layers with different W_std / b_std (nt allows it): the numpy oracle against 50-digit mpmath vectors
    (tests/golden/make_mp_layers_golden.py), NNGP and NTK, kernels and posteriors."""
    z = np.load(golden_dir / "mp_layers.npz")
    g = {k.split("/")[1]: z[k] for k in z.files if k.startswith(case + "/")}
    sw, sb, reg = tuple(g["sigma_w"]), tuple(g["sigma_b"]), float(g["diag_reg"])
    depth = len(sw)
    x, y, xt = g["x_train"], g["y_train"], g["x_test"]
    k, th = o.kernel_fn(x, None, depth, sw, sb, get="both")
    ks, ths = o.kernel_fn(xt, x, depth, sw, sb, get="both")

    def rel(a, b):
        return np.max(np.abs(a - b)) / np.max(np.abs(b))

    assert rel(k, g["K_dd"]) < 1e-14 and rel(ks, g["K_td"]) < 1e-14
    assert rel(th, g["Theta_dd"]) < 1e-8 and rel(ths, g["Theta_td"]) < 1e-8      # (theta ~ 1e-8 at duplicates)
    f = o.Fit(x, y, depth, sw, sb, diag_reg=reg)
    m, v = f.predict(xt)
    assert abs(f.lam - float(g["lam"])) < 1e-13 * f.lam
    assert rel(m, g["mean"]) < 1e-9 and rel(v, g["var"]) < 1e-8
    fn = o.FitNTK(x, y, depth, sw, sb, diag_reg=reg)
    mn, vn = fn.predict(xt)
    assert rel(mn, g["ntk_mean"]) < 1e-7 and rel(vn, g["ntk_var"]) < 1e-5

"""Staged GPU diagnostics (run under gpurun): each stage is independent and prints PASS/FAIL with
numbers, so one call localises a fault.  Not a test and not the bench; writes gpurun_out/diag.json."""
import json
import os
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "nngp-src_b200", ROOT / "oracle"):
    sys.path.insert(0, str(p))

import nngp_oracle as oracle  # noqa: E402
from nngp_b200 import _lib, synth  # noqa: E402

OUT = {}
STAGES = sys.argv[1:] or ["env", "gemm", "kernel", "potrf", "fit", "forest", "peak", "perf"]


def stage(name):
    def deco(fn):
        def run():
            if name not in STAGES:
                return
            t = time.time()
            try:
                OUT[name] = fn()
                print(f"[{name}] done in {time.time() - t:.1f}s: {json.dumps(OUT[name], default=float)[:1500]}", flush=True)
            except Exception as e:  # noqa: BLE001
                OUT[name] = {"error": repr(e)}
                print(f"[{name}] FAILED: {e!r}", flush=True)
                traceback.print_exc()
        return run
    return deco


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@stage("env")
def s_env():
    import subprocess
    r = {"cores": len(os.sched_getaffinity(0))}
    r["smi"] = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit",
                               "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    for mod in ("jax", "neural_tangents", "gpytorch"):
        try:
            __import__(mod)
            r[mod] = "present"
        except Exception as e:  # noqa: BLE001
            r[mod] = f"absent ({type(e).__name__})"
    return r


@stage("gemm")
def s_gemm():
    """depth=1 kernel == pure TMA+DMMA GEMM: K = x1 x2^T / D."""
    res = {}
    rng = np.random.default_rng(0)
    h = _lib.Handle(depth=1, stats_level=0)
    for (m, n, d) in [(5, 7, 4), (8, 8, 16), (64, 64, 16), (128, 64, 32), (130, 70, 20), (300, 200, 20),
                      (1000, 900, 130), (257, 513, 33), (2048, 2048, 128)]:
        a, b = rng.standard_normal((m, d)), rng.standard_normal((n, d))
        k = h.kernel(a, b)
        ref = a @ b.T / d
        res[f"{m}x{n}x{d}"] = relerr(k, ref)
        if res[f"{m}x{n}x{d}"] > 1e-12:
            bad = np.argwhere(np.abs(k - ref) > 1e-10 * np.max(np.abs(ref)))
            res[f"{m}x{n}x{d}_bad"] = {"count": int(len(bad)), "first": bad[:8].tolist(),
                                       "got": k[tuple(bad[0])] if len(bad) else None,
                                       "want": ref[tuple(bad[0])] if len(bad) else None}
    a = rng.standard_normal((333, 24))
    res["sym_333"] = relerr(h.kernel(a), a @ a.T / 24)
    h.close()
    return res


@stage("kernel")
def s_kernel():
    res = {}
    for depth, sw, sb in [(2, 1.0, 0.0), (3, 1.0, 0.0), (2, 1.5, 0.05), (4, 1.2, 0.1)]:
        h = _lib.Handle(depth=depth, sigma_w=sw, sigma_b=sb, stats_level=0)
        x1 = synth.encodings(700, 40, 3)
        x2 = synth.encodings(450, 40, 4)
        x1[5] = x1[4]; x1[6] = 0.0; x2[0] = x1[4]; x2[1] = 2 * x1[4]
        res[f"d{depth}_sw{sw}_sb{sb}"] = {
            "rect": relerr(h.kernel(x1, x2), oracle.kernel_fn(x1, x2, depth, sw, sb)),
            "sym": relerr(h.kernel(x1), oracle.kernel_fn(x1, None, depth, sw, sb))}
        h.close()
    return res


@stage("potrf")
def s_potrf():
    import scipy.linalg as sla
    res = {}
    rng = np.random.default_rng(1)
    h = _lib.Handle(stats_level=0)
    for n in [8, 64, 65, 128, 200, 256, 300, 513, 1000, 2048]:
        a = rng.standard_normal((n, n + 8))
        spd = a @ a.T + n * np.eye(n)
        l = h.potrf(spd)
        ref = sla.cholesky(spd, lower=True)
        res[str(n)] = relerr(l, ref)
    try:
        h.potrf(-np.eye(70))
        res["notpd"] = "NOT RAISED"
    except np.linalg.LinAlgError as e:
        res["notpd"] = str(e)[:120]
    h.close()
    return res


@stage("fit")
def s_fit():
    res = {}
    for (n, t, d, depth) in [(40, 12, 20, 2), (700, 300, 24, 2), (1500, 1000, 64, 3), (3000, 2000, 128, 2)]:
        xtr, ytr, xte, _ = synth.make_problem(n, t, d)
        h = _lib.Handle(depth=depth)
        h.fit(xtr, ytr)
        st = h.get_state()
        ref = oracle.Fit(xtr, ytr, depth)
        mean, var = h.predict(xte)
        rm, rv = ref.predict(xte)
        res[f"{n}x{t}x{d}_d{depth}"] = {
            "lam": abs(st["lambda"] - ref.lam) / ref.lam, "L": relerr(st["l"], ref.c),
            "alpha": relerr(st["alpha"], ref.alpha), "mean": relerr(mean, rm), "var": relerr(var, rv),
            "min_var_over_kss": float(np.min(rv) / np.max(rv))}
        h.close()
    return res


@stage("forest")
def s_forest():
    z = np.load(ROOT / "tests/golden/forest_xy.npz")
    xtr, ytr, xte, yte = z["x_train"], z["y_train"], z["x_test"], z["y_test"]
    h = _lib.Handle()
    t0 = time.time(); h.fit(xtr, ytr); t_fit = time.time() - t0
    t0 = time.time(); mean, var = h.predict(xte); t_pred = time.time() - t0
    t0 = time.time(); ref = oracle.Fit(xtr, ytr); t_ofit = time.time() - t0
    t0 = time.time(); rm, rv = ref.predict(xte); t_opred = time.time() - t0
    qe, rqe = oracle.q_error_stats(mean, yte), oracle.q_error_stats(rm, yte)
    return {"lambda": h.dims()[2], "oracle_lambda": ref.lam, "mean_rel": relerr(mean, rm),
            "mean_rel_pointwise_max": float(np.max(np.abs(mean - rm) / np.maximum(np.abs(rm), 1e-3))),
            "var_rel": relerr(var, rv), "var_rel_pointwise_max": float(np.max(np.abs(var - rv) / np.abs(rv))),
            "qerr": qe, "oracle_qerr": rqe, "gpu_fit_s": t_fit, "gpu_pred_s": t_pred, "cpu_fit_s": t_ofit,
            "cpu_pred_s": t_opred, "stats": h.stats()}


@stage("peak")
def s_peak():
    import torch
    h = _lib.Handle(stats_level=0)
    res = {"dmma_peak_tflops": h.dmma_peak_tflops()}
    for (m, n, k) in [(37888, 64, 4096), (37888, 64, 8192), (37888, 64, 512), (8192, 8192, 256), (8192, 8192, 128),
                      (16384, 16384, 256), (32768, 4096, 4096)]:
        ms = h.gemm_probe_ms(m, n, k, 5)
        res[f"gemm_{m}x{n}x{k}"] = {"ms": ms, "tflops": 2.0 * m * n * k / ms / 1e9}
    h.close()
    a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res["cublas_dgemm_8192_tflops"] = 2 * 8192**3 / best / 1e9
    return res


@stage("perf")
def s_perf():
    res = {}
    for (n, t, d, depth) in [(8192, 65536, 128, 2), (16384, 37888, 256, 3)]:
        xtr, ytr, xte, _ = synth.make_problem(n, t, d)
        h = _lib.Handle(depth=depth, stats_level=2)
        h.fit(xtr, ytr)      # warm-up (allocations, module load)
        h.stats_reset()
        t0 = time.time(); h.fit(xtr, ytr); t_fit = time.time() - t0
        sfit = h.stats(); h.stats_reset()
        h.predict(xte[:4096])
        h.stats_reset()
        t0 = time.time(); mean, var = h.predict(xte); t_pred = time.time() - t0
        sp = h.stats()
        res[f"N{n}_T{t}_D{d}_d{depth}"] = {
            "fit_wall_s": t_fit, "fit_ms": {k: sfit[k] for k in ("fit_gram_ms", "fit_chol_ms", "fit_solve_ms", "fit_total_ms", "h2d_ms")},
            "chol_tflops": n**3 / 3 / sfit["fit_chol_ms"] / 1e9,
            "fit_gemm_tflops": sfit["gemm_flops"] / max(sfit["gemm_ms"], 1e-9) / 1e9,
            "pred_wall_s": t_pred, "qps_wall": t / t_pred, "qps_dev": t / (sp["pred_total_ms"] / 1e3),
            "pred_ms": {k: sp[k] for k in ("pred_gram_ms", "pred_mean_ms", "pred_trsm_ms", "pred_var_ms", "pred_total_ms", "h2d_ms", "d2h_ms")},
            "pred_gemm_tflops": sp["gemm_flops"] / max(sp["gemm_ms"], 1e-9) / 1e9,
            "pred_gemm_ms": sp["gemm_ms"], "launches": sp["kernel_launches"],
            "any_nan": bool(np.isnan(mean).any() or np.isnan(var).any()), "min_var": float(np.min(var))}
        h.close()
    return res


if __name__ == "__main__":
    for fn in (s_env, s_gemm, s_kernel, s_potrf, s_fit, s_forest, s_peak, s_perf):
        fn()
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    with open(ROOT / "gpurun_out" / ("diag_" + "_".join(STAGES) + ".json"), "w") as fh:
        json.dump(OUT, fh, indent=1, default=float)
    print("DIAG COMPLETE")

"""GPU check of the int8 digit-plane product (sliced_gemm.cuh) behind ``variance_slices``.

    python tests/checks/sliced_check.py product      # nngp_sliced_product vs a numpy restatement of the same digits
    python tests/checks/sliced_check.py model [N T]  # fit + predict with variance_slices vs the FP64 path and the oracle
    python tests/checks/sliced_check.py time [N T D depth]   # timings of the two variance paths

Writes gpurun_out/sliced_<stage>.json.  Test infrastructure: imports the oracle as the checker.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "nngp-src_b200"))
sys.path.insert(0, ROOT)

from nngp_b200 import _lib  # noqa: E402


def split_rows(a, s, tri=False):
    a = np.tril(a) if tri else a
    amax = np.max(np.abs(a), axis=1)
    _, e = np.frexp(amax)
    e = np.where(amax == 0, 0, e)
    t = np.ldexp(a, (6 - e)[:, None])
    planes = []
    for _ in range(s):
        q = np.rint(t)
        planes.append(q)
        t = (t - q) * 128.0
    return planes, e


def sliced_ref(a, b, s, tri):
    pa, ea = split_rows(a, s)
    pb, eb = split_rows(b, s, tri)
    acc = np.zeros((a.shape[0], b.shape[0]))
    for g in range(s - 1, -1, -1):                      # Horner in 2^-7, as the kernel's epilogue
        c = np.zeros_like(acc)
        for p in range(g + 1):
            c += pa[p] @ pb[g - p].T
        acc = c if g == s - 1 else acc * 0.0078125 + c
    return acc * np.ldexp(1.0, ea - 6)[:, None] * np.ldexp(1.0, eb - 6)[None, :]


def describe_mismatch(v, ref):
    bad = ~np.isclose(v, ref, rtol=1e-12, atol=1e-300)
    out = {"bad_fraction": float(bad.mean())}
    if bad.any():
        r, c = np.argwhere(bad)[0]
        out["first_bad"] = [int(r), int(c), float(v[r, c]), float(ref[r, c])]
        out["bad_rows_mod128"] = sorted(set((np.argwhere(bad)[:, 0] % 128).tolist()))[:40]
        out["bad_cols_mod256"] = sorted(set((np.argwhere(bad)[:, 1] % 256).tolist()))[:40]
        out["bad_row_tiles"] = sorted(set((np.argwhere(bad)[:, 0] // 128).tolist()))[:40]
        out["bad_col_tiles"] = sorted(set((np.argwhere(bad)[:, 1] // 256).tolist()))[:40]
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = v[bad] / ref[bad]
        out["ratio_quantiles"] = [float(x) for x in np.nanquantile(ratio, [0, 0.25, 0.5, 0.75, 1])]
        out["v_zero_fraction_among_bad"] = float((v[bad] == 0).mean())
    return out


def stage_product():
    rng = np.random.default_rng(5)
    h = _lib.Handle()
    res = []
    cases = [(128, 128, 256, 0, 1), (128, 128, 256, 0, 3), (128, 512, 256, 0, 2), (200, 300, 500, 0, 7),
             (256, 1024, 1024, 1, 7), (1000, 2304, 2304, 1, 7), (4096, 4096, 4096, 1, 8), (19000, 1536, 1536, 1, 5)]
    ok_all = True
    for (m, k, n, tri, s) in cases:
        a = rng.normal(size=(m, k)) * np.exp(rng.normal(size=(m, 1)) * 3)
        b = rng.normal(size=(n, k)) * np.exp(rng.normal(size=(n, 1)) * 3)
        if tri:
            b = np.tril(b)
        t0 = time.time()
        v, rs = h.sliced_product(a, b, slices=s, lower=bool(tri), want_rowsq=True)
        dt = time.time() - t0
        ref = sliced_ref(a, b, s, bool(tri))
        exact = a @ b.T
        scale = np.max(np.abs(a), axis=1)[:, None] * np.max(np.abs(b), axis=1)[None, :] * k
        rec = {"case": [m, k, n, tri, s], "seconds": round(dt, 3),
               "max_rel_vs_digit_restatement": float(np.max(np.abs(v - ref) / np.maximum(np.abs(ref), 1e-300))),
               "max_err_vs_fp64_over_rowmax_K": float(np.max(np.abs(v - exact) / scale)),
               "rowsq_rel": float(np.max(np.abs(rs - np.einsum("ij,ij->i", v, v)) / np.einsum("ij,ij->i", v, v)))}
        rec["ok"] = bool(rec["max_rel_vs_digit_restatement"] < 1e-13 and rec["rowsq_rel"] < 1e-13)
        if not rec["ok"]:
            rec["mismatch"] = describe_mismatch(v, ref)
            ok_all = False
        print(json.dumps(rec), flush=True)
        res.append(rec)
    return {"stage": "product", "ok": ok_all, "cases": res}


def stage_model(n=2048, t=8192, d=64, depth=2, slices=(7, 8)):
    from oracle import nngp_oracle as orc
    rng = np.random.default_rng(11)
    x = rng.random((n, d))
    y = rng.normal(size=n) * 3 + 8
    xt = np.vstack([rng.random((t - t // 4, d)), x[rng.integers(0, n, t // 4)] + 1e-3 * rng.normal(size=(t // 4, d))])
    ref = _lib.Handle(depth=depth)
    ref.fit(x, y)
    m0, v0 = ref.predict(xt)
    fit = orc.Fit(x, y, depth=depth)
    sel = rng.choice(t, 512, replace=False)
    mo, vo = fit.predict(xt[sel])
    out = {"stage": "model", "N": n, "T": t, "fp64_vs_oracle_var": float(np.max(np.abs(v0[sel] - vo) / np.abs(vo))), "slices": {}}
    ok = True
    for s in slices:
        h = _lib.Handle(depth=depth, variance_slices=s)
        h.fit(x, y)
        m1, v1 = h.predict(xt)
        st = h.stats()
        rec = {"mean_bitwise": bool(np.array_equal(m0, m1)),
               "var_vs_fp64_path": float(np.max(np.abs(v1 - v0) / np.abs(v0))),
               "var_vs_oracle": float(np.max(np.abs(v1[sel] - vo) / np.abs(vo))),
               "std_vs_oracle": float(np.max(np.abs(np.sqrt(v1[sel]) - np.sqrt(vo)) / np.sqrt(vo))),
               "sliced_ms": st["sliced_ms"], "sliced_macs": st["sliced_macs"]}
        m2, v2 = h.predict(xt)
        rec["repeat_bitwise"] = bool(np.array_equal(v1, v2))
        rec["ok"] = bool(rec["mean_bitwise"] and rec["repeat_bitwise"] and rec["var_vs_oracle"] < 1e-6)
        ok = ok and rec["ok"]
        out["slices"][str(s)] = rec
        print(json.dumps({str(s): rec}), flush=True)
        h.close()
    out["ok"] = ok
    return out


def stage_time(n=8192, t=65536, d=128, depth=2, s=7):
    import torch
    rng = np.random.default_rng(3)
    x = rng.random((n, d))
    y = rng.normal(size=n) * 3 + 8
    xt = torch.from_numpy(rng.random((t, d))).cuda()
    out = {"stage": "time", "N": n, "T": t, "D": d, "depth": depth, "slices": s}
    for name, kw in (("fp64", {}), ("sliced", {"variance_slices": s})):
        h = _lib.Handle(depth=depth, stats_level=1, **kw)
        h.fit(x, y)
        mean = torch.empty(t, dtype=torch.float64, device="cuda")
        var = torch.empty(t, dtype=torch.float64, device="cuda")
        for rep in range(3):
            h.stats_reset()
            torch.cuda.synchronize()
            t0 = time.time()
            h.predict(xt, mean_out=mean, var_out=var)
            torch.cuda.synchronize()
            wall = time.time() - t0
            st = h.stats()
        rec = {"predict_wall_s": round(wall, 4), "pred_total_ms": st["pred_total_ms"], "pred_gram_ms": st["pred_gram_ms"],
               "pred_trsm_ms": st["pred_trsm_ms"], "sliced_ms": st["sliced_ms"], "queries_per_s": t / wall}
        if st["sliced_macs"]:
            rec["int8_tops"] = 2 * st["sliced_macs"] / (st["sliced_ms"] * 1e-3) / 1e12
            rec["fp64_equivalent_tflops"] = float(t) * n * n / (st["sliced_ms"] * 1e-3) / 1e12
        out[name] = rec
        out[name + "_var_sample"] = var[:4].cpu().tolist()
        print(json.dumps({name: rec}), flush=True)
        h.close()
        del h
        torch.cuda.empty_cache()
    return out


def main():
    stage = sys.argv[1] if len(sys.argv) > 1 else "product"
    args = [int(v) for v in sys.argv[2:]]
    res = {"product": stage_product, "model": stage_model, "time": stage_time}[stage](*args)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = stage + ("_" + "_".join(str(a) for a in args) if args else "")
    with open(os.path.join(ROOT, "gpurun_out", f"sliced_{tag}.json"), "w") as f:
        json.dump(res, f, indent=1)
    print("RESULT", stage, "ok" if res.get("ok", True) else "FAILED")
    return 0 if res.get("ok", True) else 1


if __name__ == "__main__":
    sys.exit(main())

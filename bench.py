#!/usr/bin/env python
"""bench.py -- headline benchmark of the NNGP hot path on B200.

Metric (BASELINE.json): predicted queries/s (posterior mean + variance) at 1/2/4/8 B200, plus train-fit
seconds and % of measured FP64 tensor (DMMA) peak.  A "step" is one prediction pass over one batch of
synthetic test queries with the fitted model resident in HBM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5] [--impl reference]

N>1 is launched by torchrun (one rank per GPU, NCCL): rank 0 fits, the fitted state is broadcast once, each
rank predicts its own shard of test rows (weak scaling: per-GPU rows fixed), no data-path collective.
Rank 0 prints ONE JSON line.  `--impl reference` times the CPU restatement of the reference path
(oracle/nngp_oracle.py: numpy/scipy -> OpenBLAS/LAPACK, the library class jaxlib calls) on the box's host
cores; jax / neural-tangents are not installable, so no `oracle/_ref` exists.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "nngp-src_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

WORKLOADS = {
    # name: (n_train, n_test_per_gpu, dim, depth, join_dims, description)
    "c2": (8192, 65536, 128, 2, 0, "C2 synthetic single-table: 8k train / 64k test per GPU, 128-dim, depth-2 NNGP"),
    "c3": (32768, 131072, 256, 3, 0, "C3 synthetic: 32k train / 128k test per GPU (1M/8), 256-dim, depth-3 NNGP"),
    "c5": (16384, 524288, 512, 3, 96, "C5 multi-join: 16k train / 512k test per GPU (4M/8), 512-dim, depth-3 NNGP"),
}
METRIC, UNIT = "predicted_queries_per_sec", "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit32k", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- reference arm
def cpu_sample(n, d, depth, join_dims, rows):
    """Fit the oracle at the workload's N, then time prediction of `rows` test rows -> queries/s."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import nngp_oracle as oracle
    from nngp_b200 import synth
    xtr, ytr, xte, _ = synth.make_problem(n, rows, d, join_dims=join_dims)
    t0 = time.perf_counter()
    fit = oracle.Fit(xtr, ytr, depth)
    t_fit = time.perf_counter() - t0
    return fit, xte, t_fit


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, _t, d, depth, jd, desc = WORKLOADS[args.workload]
    rows = 2048 if n <= 8192 else 512
    cores = len(os.sched_getaffinity(0))
    fit, xte, t_fit = cpu_sample(n, d, depth, jd, rows)
    for _ in range(args.warmup):
        fit.predict(xte[:256])
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        fit.predict(xte)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    val = rows / sec
    sample = f"oracle fit at N={n} ({t_fit:.1f}s, untimed) then predict mean+var of {rows} test rows per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n_train": n, "dim": d, "depth": depth, "rows_per_step": rows},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "fit_seconds": t_fit},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from nngp_b200 import _lib, batch, predict, runtime, stax, synth
    from nngp_b200 import dist as ndist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun (--nproc-per-node {args.gpus}); WORLD_SIZE is 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the nngp_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    runtime.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, t_rank, d, depth, jd, desc = WORKLOADS[args.workload]
    xtr, ytr = synth.encodings(n, d, 1, join_dims=jd), None
    ytr = synth.labels(xtr, jd)
    xte_host = synth.encodings(t_rank, d, 2 + rank, join_dims=jd)

    # the reference-facing call chain (train.py:161-172); the engine below is the C-ABI handle behind predict_fn
    _, _, kernel_fn = stax.serial(*([stax.Dense(512)] + [l for _ in range(depth - 1) for l in (stax.Relu(), stax.Dense(1))]))
    kernel_fn = batch.batch(kernel_fn, device_count=0, batch_size=0)
    runtime.set_stats_level(2)
    predict_fn = predict.gradient_descent_mse_ensemble(kernel_fn, xtr, ytr, diag_reg=1e-3)

    fit_info = {}
    if rank == 0:
        h = predict_fn.engine()            # fit #1 (allocations, module load) -- warm-up
        h.stats_reset()
        t0 = time.perf_counter()
        h.fit(xtr, ytr)                    # fit #2, timed
        fit_wall = time.perf_counter() - t0
        s = h.stats()
        fit_info = {"n_train": n, "dim": d, "depth": depth, "seconds_device": s["fit_total_ms"] / 1e3,
                    "seconds_wall_incl_h2d": fit_wall, "gram_ms": s["fit_gram_ms"], "chol_ms": s["fit_chol_ms"],
                    "solve_ms": s["fit_solve_ms"], "chol_tflops": n**3 / 3 / max(s["fit_chol_ms"], 1e-9) / 1e9}
    else:
        h = runtime.new_handle(kernel_fn.spec, diag_reg=1e-3)
    bcast_ms = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ndist.broadcast_fit(h, src=0)
        torch.cuda.synchronize()
        dist.barrier()
        bcast_ms = (time.perf_counter() - t0) * 1e3

    stream = torch.cuda.ExternalStream(h.stream, device=torch.device("cuda", local))
    xte_dev = torch.from_numpy(xte_host).cuda()
    mean_dev = torch.empty(t_rank, dtype=torch.float64, device="cuda")
    var_dev = torch.empty(t_rank, dtype=torch.float64, device="cuda")
    xte_pinned = torch.from_numpy(xte_host).pin_memory()
    mean_pin = torch.empty(t_rank, dtype=torch.float64).pin_memory()
    var_pin = torch.empty(t_rank, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():      # inputs/outputs already in HBM: the C-ABI call on device pointers
        h.predict(xte_dev, want_var=True, mean_out=mean_dev, var_out=var_dev)

    def step_e2e():           # host buffers in, host results out, through the same C-ABI entry point
        h.predict(xte_pinned.numpy(), want_var=True, mean_out=mean_pin.numpy(), var_out=var_pin.numpy())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    h.stats_reset()
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    st = h.stats()

    for _ in range(2):
        step_e2e()
    h.stats_reset()
    ms_e2e = timed(step_e2e, args.steps)
    st_e2e = h.stats()

    # parity spot-check on the way out is tests' job; here only sanity
    m = mean_dev[:1024].cpu().numpy()
    if not np.all(np.isfinite(m)):
        raise SystemExit("bench.py: non-finite predictions")

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    queries = world * t_rank * args.steps
    value = queries / (ms_total / 1e3)
    e2e_value = queries / (ms_e2e / 1e3)
    peak_dmma = h.dmma_peak_tflops()
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    peaks = json.loads(peaks_file.read_text()) if peaks_file.exists() else {}
    achieved = st["gemm_flops"] / max(st["gemm_ms"], 1e-9) / 1e9
    ncu_file = ROOT / "profiles" / "ncu_dominant_kernel.json"
    traffic = json.loads(ncu_file.read_text()).get("dram_bytes_per_launch") if ncu_file.exists() else None
    flops_per_query = float(n) * n + 2.0 * n * d + 4.0 * n
    roofline = {
        "bound": "tensor", "kernel": "trsm_fused_kernel (persistent TMA + DMMA.8x8x4 FP64 triangular solve V = K_* L^-T with fused variance)",
        "achieved": achieved, "peak": peak_dmma, "unit": "TFLOP/s", "frac": achieved / peak_dmma,
        "traffic": traffic,
        "peak_source": "FP64 tensor (DMMA) issue-rate microbenchmark measured in this run (burst); MEASURED_PEAKS.json "
                       "has no FP64 entry (bf16 %.0f TF/s, HBM %.0f GB/s are not the bound of an FP64 kernel)" % (
                           peaks.get("bf16_tflops", float("nan")), peaks.get("hbm_gbs", float("nan"))),
        "launches": st["gemm_launches"], "avg_launch_ms": st["gemm_ms"] / max(st["gemm_launches"], 1),
        "algorithmic_flops_per_launch": st["gemm_flops"] / max(st["gemm_launches"], 1),
        "share_of_step": st["gemm_ms"] / max(st["pred_total_ms"], 1e-9),
        "whole_step_tflops_per_gpu": flops_per_query * t_rank * args.steps / (ms_total / 1e3) / 1e12,
        "whole_step_frac_of_peak": flops_per_query * t_rank * args.steps / (ms_total / 1e3) / 1e12 / peak_dmma,
    }
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n_train": n, "test_rows_per_gpu": t_rank, "dim": d, "depth": depth,
                   "diag_reg": 1e-3, "parallelism": f"fit on rank 0 + broadcast, test rows sharded x{world}",
                   "l2": "inputs larger than L2: per step the factor L (%.0f MB) and the K_* block (%.0f MB) stream "
                         "through a 126 MB L2" % (n * n * 8 / 1e6, t_rank * n * 8 / 1e6)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": st_e2e["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st_e2e["d2h_bytes"] // args.steps,
                "api": "nngp_predict (C ABI) on pinned host buffers == predict_fn(x_test=...) of the neural-tangents mirror"},
        "gpu_launches": st["kernel_launches"],
        "roofline": roofline,
        "stage_ms_per_step": {k: st[k] / args.steps for k in ("pred_gram_ms", "pred_mean_ms", "pred_trsm_ms", "pred_var_ms", "pred_total_ms")},
        "fit": fit_info,
    }
    if bcast_ms is not None:
        out["fit_broadcast_ms"] = bcast_ms

    if world == 1 and not args.no_fit32k:
        try:
            h.close()
            del xte_dev, mean_dev, var_dev
            torch.cuda.empty_cache()
            n32, d32, depth32 = 32768, 256, 3
            x32 = synth.encodings(n32, d32, 1)
            y32 = synth.labels(x32)
            h32 = _lib.Handle(depth=depth32, diag_reg=1e-3, device=local, stats_level=2)
            h32.fit(x32, y32)
            h32.stats_reset()
            h32.fit(x32, y32)
            s = h32.stats()
            out["fit_n32k"] = {"n_train": n32, "dim": d32, "depth": depth32, "seconds_device": s["fit_total_ms"] / 1e3,
                               "gram_ms": s["fit_gram_ms"], "chol_ms": s["fit_chol_ms"], "solve_ms": s["fit_solve_ms"],
                               "chol_tflops": n32**3 / 3 / s["fit_chol_ms"] / 1e9,
                               "chol_frac_of_fp64_peak": n32**3 / 3 / s["fit_chol_ms"] / 1e9 / peak_dmma,
                               "gemm_kernel_tflops": s["gemm_flops"] / max(s["gemm_ms"], 1e-9) / 1e9}
            h32.close()
        except Exception as e:  # noqa: BLE001
            out["fit_n32k"] = {"error": repr(e)}

    if world == 1 and not args.no_cpu_baseline:
        rows = 2048 if n <= 8192 else 512
        t0 = time.perf_counter()
        fit, xs, t_fit = cpu_sample(n, d, depth, jd, rows)
        t1 = time.perf_counter()
        fit.predict(xs)
        t_pred = time.perf_counter() - t1
        out["cpu_baseline"] = {"value": rows / t_pred, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                               "sample": f"oracle (numpy/scipy, OpenBLAS all cores): fit N={n} in {t_fit:.1f}s, then predict "
                                         f"mean+var of {rows} test rows in {t_pred:.2f}s",
                               "fit_seconds": t_fit}
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    # stdout carries exactly ONE line (rank 0's JSON): library banners (e.g. "NCCL version ...") that are
    # written to fd 1 during the run are diverted to stderr, and fd 1 is restored for the final print.
    sys.stdout.flush()
    _saved_fd = os.dup(1)
    os.dup2(2, 1)
    _real_print = print

    def print(*args, **kw):  # noqa: A001
        sys.stdout.flush()
        os.dup2(_saved_fd, 1)
        _real_print(*args, **kw)
        sys.stdout.flush()
        os.dup2(2, 1)

    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)

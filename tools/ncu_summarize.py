"""Summarise ncu outputs into small text/JSON files for profiles/ (run here, no GPU needed).
  python tools/ncu_summarize.py launches <launches.csv> <out.txt>
  python tools/ncu_summarize.py full <report.ncu-rep> <out.json>
"""
import collections
import csv
import json
import re
import subprocess
import sys


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none : {n} launches, {tot:.3f} ms total "
                 f"(cold-cache, serialised: compare SHARES)\n")
        fh.write(f"{'kernel':44s} {'launches':>8s} {'total_ms':>10s} {'share':>7s} {'avg_us':>10s} {'max_us':>10s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"{k:44s} {v[0]:8d} {v[1]:10.3f} {v[1] / tot:7.3f} {1e3 * v[1] / v[0]:10.2f} {1e3 * v[2]:10.2f}\n")
    print(open(out).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.sum", "sm__ops_path_tensor_src_fp64.sum.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]][:80], "grid": r[idx.get("Grid Size", 0)], "block": r[idx.get("Block Size", 0)]}
        for w in WANT:
            if w in idx:
                d[w] = f"{r[idx[w]]} {units[idx[w]]}".strip()
        for h in hdr:      # tcgen05 kernels: every tensor / TMEM / L2-fabric counter the capture holds
            if re.search(r"pipe_tensor|tmem|utc|lts__t_sectors_srcunit_tex|l1tex__m_xbar2l1tex_read_bytes|lts__throughput|sm__cycles_active.avg$|clocks", h) and h not in d:
                val = r[idx[h]]
                if val not in ("", "0", "n/a"):
                    d[h] = f"{val} {units[idx[h]]}".strip()
        for h in hdr:
            if "issue_stalled" in h and h.endswith("_per_warp_active.pct") or ("warp_issue_stalled" in h and h.endswith(".pct")):
                try:
                    if float(r[idx[h]]) >= 3.0:
                        d.setdefault("stalls_pct", {})[h.split("warp_issue_stalled_")[-1].split("_per_warp")[0]] = float(r[idx[h]])
                except ValueError:
                    pass
        for h in hdr:      # warp stall reasons, in issue slots per issued instruction
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    if float(r[idx[h]]) >= 0.2:
                        name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                        d.setdefault("stall_slots_per_issue", {})[name] = round(float(r[idx[h]]), 3)
                except ValueError:
                    pass
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

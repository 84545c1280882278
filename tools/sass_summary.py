"""Per-kernel SASS / resource summary of libnngp_b200.so (runs here: cuobjdump needs no GPU).
    python tools/sass_summary.py [out.txt]
For every kernel: registers, stack, local, static shared memory, and a histogram of the instructions that show what
the kernel is built from -- DMMA (FP64 tensor core), UTMALDG (TMA tensor loads), SYNCS (mbarrier), LDS/STS, DFMA/DMUL/
DADD (FP64 CUDA cores), MUFU (sqrt / reciprocal seeds), and UTC* / LDTM / STTM (tcgen05: only sliced_gemm_kernel, the
int8 digit-plane product -- tcgen05 has no FP64 kind: DESIGN.md sections 3 and 5.10)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "nngp-src_b200" / "libnngp_b200.so"
TOOL = "/usr/local/cuda/bin/cuobjdump"
KEYS = ["DMMA", "UTMALDG", "SYNCS", "LDS", "STS", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "ATOM", "BAR", "UTC", "LDTM", "STTM"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", n).replace("void ", "").replace("nngp::", "") for n in out]


def main():
    res = subprocess.run([TOOL, "--dump-resource-usage", str(LIB)], capture_output=True, text=True).stdout
    usage = {m[0]: m[1:] for m in re.findall(r"Function (\S+):\s*\n?\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res)}
    sass = subprocess.run([TOOL, "-sass", str(LIB)], capture_output=True, text=True).stdout
    hist, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            hist[cur]["total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    hist[cur][k] += 1
    names = list(hist)
    pretty = dict(zip(names, demangle(names)))
    build = subprocess.run([sys.executable, str(ROOT / "nngp-src_b200" / "nngp_b200" / "_build.py"), "--hash"],
                           capture_output=True, text=True).stdout.strip()
    nvcc = subprocess.run(["/usr/local/cuda/bin/nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2]
    lines = [f"# libnngp_b200.so  build id {build}  ({nvcc}); cuobjdump -sass / --dump-resource-usage, sm_100a",
             "# FP64 kernels: tcgen05.mma has no FP64 kind, their tensor route on sm_100a is mma.sync.m8n8k4.f64 -> DMMA.8x8x4,",
             "# fed by TMA (UTMALDG) through mbarriers (SYNCS).  sliced_gemm_kernel (variance_slices) is the tcgen05 kernel:",
             "# UTCIMMA (kind::i8 MMA), UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM alloc), LDTM (tcgen05.ld), counted under UTC / LDTM.",
             f"{'kernel':44s} {'REG':>4s} {'STACK':>5s} {'SMEM':>6s} {'LOCAL':>5s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS)]
    tot = collections.Counter()
    for n in names:
        u = usage.get(n, ("?", "?", "?", "?"))
        h = hist[n]
        tot.update(h)
        lines.append(f"{pretty[n][:44]:44s} {u[0]:>4s} {u[1]:>5s} {u[2]:>6s} {u[3]:>5s} {h['total']:6d} " + " ".join(f"{h[k]:7d}" for k in KEYS))
    lines.append(f"{'TOTAL':44s} {'':>4s} {'':>5s} {'':>6s} {'':>5s} {tot['total']:6d} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(text)
    print(text)


if __name__ == "__main__":
    main()

#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/bench_active.py > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -c 1200 gpurun_out/r02_bench_c4.json
timeout 300 python tools/bench_estimator.py > gpurun_out/r02_bench_estimator.json 2> gpurun_out/r02_bench_estimator.err; tail -c 600 gpurun_out/r02_bench_estimator.json
timeout 300 python tests/checks/illcond_report.py > gpurun_out/r02_illcond.json 2> gpurun_out/r02_illcond.err; cat gpurun_out/r02_illcond.json
for w in 512 768 1024; do echo "--- W=$w at 32768"; NNGP_CHOL_W=$w timeout 120 python tools/fit_once.py 32768 256 3 2; done
NNGP_CHOL_LOOKAHEAD=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 2 -c 1 -o gpurun_out/r02_syrk_32k python tools/fit_once.py 32768 256 3 0 > gpurun_out/ncu_syrk.log 2>&1; tail -2 gpurun_out/ncu_syrk.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_c3.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench_c3.log 2>&1
python tools/ncu_summarize.py launches gpurun_out/r02_launches_bench_c3.csv gpurun_out/r02_launches_bench_c3.txt | head -20
echo "--- gram (integer compares, hoisted diag updates)"; python tools/gram_once.py 16384 8192 128 2 5; python tools/gram_once.py 16384 32768 256 3 3; python tools/gram_once.py 16384 16384 512 3 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2

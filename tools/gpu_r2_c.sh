#!/usr/bin/env bash
# round-2 GPU job C (2 GPUs): gpu tests again, latency sweep, fit launch breakdown (ncu), C2 bench (Gram kernel)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40 > gpurun_out/r02_gputests_c.log
tail -12 gpurun_out/r02_gputests_c.log
python tools/latency_sweep.py > gpurun_out/r02_latency_sweep.json 2> gpurun_out/r02_latency_sweep.err
head -50 gpurun_out/r02_latency_sweep.json; tail -3 gpurun_out/r02_latency_sweep.err
python tools/fit_once.py 8192 128 2 5
NNGP_PANEL_SOLVE=fma python tools/fit_once.py 8192 128 2 5
NNGP_CHOL_LOOKAHEAD=0 python tools/fit_once.py 8192 128 2 5
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fit8k.csv python tools/fit_once.py 8192 128 2 1 > gpurun_out/ncu_fit.log 2>&1
python tools/ncu_summarize.py launches gpurun_out/r02_launches_fit8k.csv gpurun_out/r02_launches_fit8k.txt | head -30
python bench.py --workload c2 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_c2_c.json 2> gpurun_out/r02_bench_c2_c.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_c.json')); print('C2', d['value'], d['roofline']['frac'], d['roofline']['gram_kernel'], d['fit'])"
NNGP_GRAM=tiles python bench.py --workload c2 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_c2_c_tiles.json 2> gpurun_out/r02_bench_c2_c_tiles.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2_c_tiles.json')); print('C2 tiles', d['value'], d['roofline']['gram_kernel'], d['fit'])"
python tools/fit_once.py 32768 256 3 1

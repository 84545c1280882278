#!/usr/bin/env bash
# round-2 GPU job A: env probe, gpu tests (1 GPU), short C3 bench
mkdir -p gpurun_out
python -c "
import json,importlib
out={}
for m in ('jax','jaxlib','neural_tangents','gpytorch'):
    try: importlib.import_module(m); out[m]='importable'
    except Exception as e: out[m]=repr(e)
import os,psutil
out['cores']=len(os.sched_getaffinity(0)); out['ram_gb']=psutil.virtual_memory().total/2**30
json.dump(out,open('gpurun_out/r02_env.json','w'),indent=1); print(out)
"
nvidia-smi -L
python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 > gpurun_out/r02_gputests_a.log
tail -5 gpurun_out/r02_gputests_a.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c3_a.json 2> gpurun_out/r02_bench_c3_a.err
tail -c 3000 gpurun_out/r02_bench_c3_a.json
tail -5 gpurun_out/r02_bench_c3_a.err

#!/usr/bin/env bash
# round-2 GPU job D (1 GPU): tests, fit times, Gram variants, ncu captures (Gram epilogue, C3 solve)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -15 > gpurun_out/r02_gputests_d.log
tail -8 gpurun_out/r02_gputests_d.log
for n in 2048 4096 8192 16384; do python tools/fit_once.py $n 128 2 5; done
echo "--- gram ILP 4 (default)"; python tools/gram_once.py 16384 8192 128 2 5; python tools/gram_once.py 16384 32768 256 3 3; python tools/gram_once.py 16384 16384 512 3 3
echo "--- gram ILP 2"; NNGP_B200_LIB=build/libnngp_ilp2.so python tools/gram_once.py 16384 8192 128 2 5; NNGP_B200_LIB=build/libnngp_ilp2.so python tools/gram_once.py 16384 32768 256 3 3
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -c 1 -s 1 -o gpurun_out/r02_gram_c2 python tools/gram_once.py 16384 8192 128 2 1 > gpurun_out/ncu_gram.log 2>&1
tail -2 gpurun_out/ncu_gram.log
python tools/latency_sweep.py > gpurun_out/r02_latency_sweep.json 2> gpurun_out/r02_latency_sweep.err
python -c "
import json; d=json.load(open('gpurun_out/r02_latency_sweep.json')); print(d['latency_mode_ms'], d['latency_mode_fit'])"

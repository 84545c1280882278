#!/usr/bin/env bash
# round-2 GPU job B (2 GPUs): all gpu tests incl. multi-GPU / NCCL / latency, short C3 bench at N=2, latency sweep
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -60 > gpurun_out/r02_gputests_b.log
tail -15 gpurun_out/r02_gputests_b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_c3_g2.json 2> gpurun_out/r02_bench_c3_g2.err
tail -c 2500 gpurun_out/r02_bench_c3_g2.json; tail -5 gpurun_out/r02_bench_c3_g2.err
python tools/latency_sweep.py > gpurun_out/r02_latency_sweep.json 2> gpurun_out/r02_latency_sweep.err
cat gpurun_out/r02_latency_sweep.json | head -60; tail -3 gpurun_out/r02_latency_sweep.err

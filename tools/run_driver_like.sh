#!/usr/bin/env bash
# What the driver runs at round end: reference arm, then our arm, same flags.
#   bash tools/run_driver_like.sh [N] [steps] [warmup]
N=${1:-1}; K=${2:-20}; W=${3:-5}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577"; fi
SECONDS=0
$L bench.py --impl reference --gpus $N --steps $K --warmup $W > gpurun_out/r02_ref_g$N.json 2> gpurun_out/r02_ref_g$N.err
echo "reference arm: ${SECONDS}s"; tail -c 1500 gpurun_out/r02_ref_g$N.json
SECONDS=0
$L bench.py --gpus $N --steps $K --warmup $W > gpurun_out/r02_bench_g$N.json 2> gpurun_out/r02_bench_g$N.err
echo "b200 arm: ${SECONDS}s"; tail -c 4500 gpurun_out/r02_bench_g$N.json; tail -3 gpurun_out/r02_bench_g$N.err

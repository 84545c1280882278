#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
export PYTHONPATH=/root/repo/nngp-src_b200
timeout 300 python tests/checks/sliced_check.py product > gpurun_out/sliced_product.log 2>&1; echo "product rc=$?"
tail -12 gpurun_out/sliced_product.log
timeout 300 python tests/checks/sliced_check.py model > gpurun_out/sliced_model.log 2>&1; echo "model rc=$?"
tail -6 gpurun_out/sliced_model.log
timeout 300 python tests/checks/sliced_check.py time 8192 65536 128 2 > gpurun_out/sliced_time.log 2>&1; echo "time rc=$?"
tail -4 gpurun_out/sliced_time.log
timeout 600 python tests/checks/sliced_check.py time 32768 131072 256 3 > gpurun_out/sliced_time_c3.log 2>&1; echo "time c3 rc=$?"
tail -4 gpurun_out/sliced_time_c3.log

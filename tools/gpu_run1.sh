#!/bin/bash
# first GPU session: hardware probes + parity tests, each under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout -s KILL 120 tools/microbench > gpurun_out/microbench.log 2>&1; echo "microbench rc=$?" >> gpurun_out/microbench.log
for t in ctc decode optim gemm lstm; do
  timeout -s KILL 600 python -m pytest tests/test_${t}_gpu.py -x -q -m gpu > gpurun_out/test_${t}.log 2>&1
  echo "rc=$?" >> gpurun_out/test_${t}.log
  tail -5 gpurun_out/test_${t}.log
done

#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 180 tools/microbench > gpurun_out/microbench.log 2>&1; echo "microbench rc=$?" >> gpurun_out/microbench.log
for t in ctc model; do
  timeout -s KILL 600 python -m pytest tests/test_${t}_gpu.py -x -q -m gpu > gpurun_out/test_${t}.log 2>&1
  echo "rc=$?" >> gpurun_out/test_${t}.log
  tail -5 gpurun_out/test_${t}.log
done
timeout -s KILL 300 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/bench_gemm.log
cat gpurun_out/bench_gemm.log

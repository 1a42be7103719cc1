#!/bin/bash
# 2-GPU validation: NCCL through the C ABI, bucketed exchange, bench at N=2 (weak + strong), tower test
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dp_check.py > gpurun_out/r2_dp12.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dp12.log; tail -12 gpurun_out/r2_dp12.log
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench12_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_bench12_n2.log; tail -4 gpurun_out/r2_bench12_n2.log
B2_BENCH_SCALING=strong timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench12_n2s.log 2>&1; echo "rc=$?" >> gpurun_out/r2_bench12_n2s.log; tail -4 gpurun_out/r2_bench12_n2s.log
B2_BENCH_COMM=torch timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench12_n2t.log 2>&1; echo "rc=$?" >> gpurun_out/r2_bench12_n2t.log; tail -4 gpurun_out/r2_bench12_n2t.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench12_n1.log 2>&1; tail -2 gpurun_out/r2_bench12_n1.log
timeout 600 python -m pytest tests/test_towers_gpu.py -q > gpurun_out/r2_test12.log 2>&1; tail -5 gpurun_out/r2_test12.log

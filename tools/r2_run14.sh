#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py -q -k "wide or grid" > gpurun_out/r2_test14.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test14.log; tail -15 gpurun_out/r2_test11.log
timeout 600 python tools/bench_wide.py > gpurun_out/r2_wide14.log 2>&1; cat gpurun_out/r2_wide11.log
timeout 900 python tools/bench_configs.py cfg4 32 1500 > gpurun_out/r2_cfg14.log 2>&1; tail -20 gpurun_out/r2_cfg11.log

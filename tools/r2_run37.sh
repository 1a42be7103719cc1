#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_vgg_gpu.py tests/test_lstm_gpu.py -q -x > gpurun_out/r2_test37.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test37.log; tail -4 gpurun_out/r2_test37.log
B2_WIDE_ONLY=1 timeout 600 python tools/bench_wide.py > gpurun_out/r2_wide37.log 2>&1; cat gpurun_out/r2_wide37.log
timeout 900 python tools/bench_configs.py cfg4 32 1500 > gpurun_out/r2_cfg37.log 2>&1; tail -2 gpurun_out/r2_cfg37.log

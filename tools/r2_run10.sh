#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py -q -k "wide or grid" > gpurun_out/r2_test10.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test10.log; tail -15 gpurun_out/r2_test10.log
timeout 600 python tools/bench_wide.py > gpurun_out/r2_wide10.log 2>&1; cat gpurun_out/r2_wide10.log
timeout 300 python -m pytest tests/test_vgg_gpu.py -q > gpurun_out/r2_test10b.log 2>&1; tail -3 gpurun_out/r2_test10b.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ctc_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/r2_test23.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test23.log; tail -5 gpurun_out/r2_test23.log
timeout 300 python tools/bench_ctc.py > gpurun_out/r2_ctc23.log 2>&1; cat gpurun_out/r2_ctc23.log

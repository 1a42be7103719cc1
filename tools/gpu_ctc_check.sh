#!/bin/bash
# CTC kernels: parity tests + CUDA-event timings at the BASELINE shapes (run under gpurun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ctc_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -3
timeout 300 python tools/bench_ctc.py 2>&1 | tee gpurun_out/ctc_bench.log

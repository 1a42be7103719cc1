#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py -q -k "wide or grid" > gpurun_out/r2_test36.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test36.log; tail -4 gpurun_out/r2_test36.log
B2_WIDE_ONLY=1 timeout 600 python tools/bench_wide.py > gpurun_out/r2_wide36.log 2>&1; cat gpurun_out/r2_wide36.log
B2ASR_LIB=$PWD/tensorflow_end2end_speech_recognition_b200/libb2asr_timing.so B2_REC_DBG=1 B2_WIDE_ONLY=1 WIDE_T=1500 timeout 300 python tools/bench_wide.py 2>&1 | grep -E "wide fwd dbg" | tail -2
timeout 900 python tools/bench_configs.py cfg4 32 1500 > gpurun_out/r2_cfg36.log 2>&1; tail -2 gpurun_out/r2_cfg36.log

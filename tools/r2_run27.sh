#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./tools/microbench7 > gpurun_out/r2_clock27.log 2>&1; cat gpurun_out/r2_clock27.log
B2ASR_LIB=$PWD/tensorflow_end2end_speech_recognition_b200/libb2asr_timing.so B2_REC_DBG=1 timeout 300 python tools/bench_rec.py --quick 2>&1 | grep -E "dbg3|dbg4|fwd only" | head -8 > gpurun_out/r2_clock27b.log; cat gpurun_out/r2_clock27b.log

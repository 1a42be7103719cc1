#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B2ASR_LIB=$PWD/tensorflow_end2end_speech_recognition_b200/libb2asr_timing.so B2_REC_DBG=1 B2_WIDE_ONLY=1 WIDE_T=1500 timeout 300 python tools/bench_wide.py > gpurun_out/r2_wide35.log 2>&1; grep -E "wide fwd dbg|fwd only" gpurun_out/r2_wide35.log | tail -5

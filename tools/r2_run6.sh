#!/bin/bash
# round-2 visit 6: new components (multitask CTC, TF beam search, PER/CER/WER, towers), BPTT phase timers, full suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multitask_gpu.py tests/test_decode_gpu.py tests/test_towers_gpu.py tests/test_optim_gpu.py tests/test_compat_gpu.py -q > gpurun_out/r2_test6.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test6.log; tail -25 gpurun_out/r2_test6.log
B2ASR_LIB=$PWD/tensorflow_end2end_speech_recognition_b200/libb2asr_timing.so B2_REC_DBG=1 timeout 120 python tools/bench_rec.py --quick 2>&1 | grep "bwd dbg" | sort | uniq -c | sort -rn | head -3 > gpurun_out/r2_rec6_timing.log; cat gpurun_out/r2_rec6_timing.log
B2ASR_LIB=$PWD/tensorflow_end2end_speech_recognition_b200/libb2asr_timing.so B2_REC_NCHAIN=1 B2_REC_DBG=1 timeout 120 python tools/bench_rec.py --quick 2>&1 | grep "bwd dbg" | sort | uniq -c | sort -rn | head -3 > gpurun_out/r2_rec6_timing_n1.log; cat gpurun_out/r2_rec6_timing_n1.log
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_multitask_gpu.py --deselect tests/test_decode_gpu.py --deselect tests/test_towers_gpu.py --deselect tests/test_optim_gpu.py --deselect tests/test_compat_gpu.py > gpurun_out/r2_test6_full.log 2>&1
echo "full rc=$?" >> gpurun_out/r2_test6_full.log; tail -6 gpurun_out/r2_test6_full.log

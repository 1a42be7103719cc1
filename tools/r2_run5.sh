#!/bin/bash
# round-2 visit 5: per-layer backward scratch in the reserve (side-stream race fix), hoisted BPTT math, phase timers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_model_gpu.py tests/test_lstm_gpu.py tests/test_seq2seq_gpu.py tests/test_towers_gpu.py tests/test_compat_gpu.py -x -q > gpurun_out/r2_test5.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test5.log; tail -4 gpurun_out/r2_test5.log
B2_SIDE_STREAM=0 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_lstm_rec_tc_gpu.py -x -q > gpurun_out/r2_test5_ss0.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test5_ss0.log; tail -2 gpurun_out/r2_test5_ss0.log
timeout 120 python tools/bench_rec.py --quick > gpurun_out/r2_rec5.log 2>&1; cat gpurun_out/r2_rec5.log
B2ASR_LIB=$PWD/tensorflow_end2end_speech_recognition_b200/libb2asr_timing.so B2_REC_DBG=1 timeout 120 python tools/bench_rec.py --quick 2>&1 | sort | uniq -c | sort -rn | head -8 > gpurun_out/r2_rec5_timing.log; cat gpurun_out/r2_rec5_timing.log
B2ASR_LIB=$PWD/tensorflow_end2end_speech_recognition_b200/libb2asr_timing.so B2_REC_NCHAIN=1 B2_REC_DBG=1 timeout 120 python tools/bench_rec.py --quick 2>&1 | sort | uniq -c | sort -rn | head -8 > gpurun_out/r2_rec5_timing_n1.log; cat gpurun_out/r2_rec5_timing_n1.log
for kp in 0.8 1; do for ss in 0 1; do
  B2_BENCH_KEEP_PROB=$kp B2_SIDE_STREAM=$ss timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench5_kp${kp}_ss$ss.json 2> gpurun_out/r2_bench5.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench5_kp${kp}_ss$ss.json").read().strip().splitlines()[-1])
    print("keep_prob=$kp side_stream=$ss ms/step %.3f e2e ms %.3f clocks %s loss %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"], d["e2e"]["loss"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench5.err").read()[-1500:])
PY
done; done

#!/bin/bash
# round-2 visit 3: BPTT kernel with warp-specialised register budget; wait-scope variants; gate-team size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_model_gpu.py tests/test_lstm_gpu.py tests/test_seq2seq_gpu.py -x -q > gpurun_out/r2_test3.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test3.log; tail -3 gpurun_out/r2_test3.log
B2_REC_GW=4 timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/r2_test3_gw4.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test3_gw4.log; tail -3 gpurun_out/r2_test3_gw4.log
for w in 1 2; do
B2_REC_WAIT=$w timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/r2_test3_wait$w.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test3_wait$w.log; tail -3 gpurun_out/r2_test3_wait$w.log
done
: > gpurun_out/r2_rec3.log
for gw in 4 8; do for w in 0 1 2; do
  B2_REC_GW=$gw B2_REC_WAIT=$w timeout 120 python tools/bench_rec.py --quick >> gpurun_out/r2_rec3.log 2>&1
done; done
cat gpurun_out/r2_rec3.log

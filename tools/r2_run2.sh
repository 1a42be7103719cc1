#!/bin/bash
# round-2 visit 2: gate-team variants of the recurrence kernels (GW=4 vs 8), correctness + timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for gw in 4 8; do
  B2_REC_GW=$gw timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_model_gpu.py tests/test_lstm_gpu.py -x -q > gpurun_out/r2_test_gw$gw.log 2>&1
  echo "gw=$gw rc=$?" >> gpurun_out/r2_test_gw$gw.log
  tail -3 gpurun_out/r2_test_gw$gw.log
  B2_REC_GW=$gw timeout 300 python tools/bench_rec.py --bwd > gpurun_out/r2_rec_gw$gw.log 2>&1
  cat gpurun_out/r2_rec_gw$gw.log
done
timeout 900 python -m pytest tests/test_towers_gpu.py tests/test_optim_gpu.py tests/test_compat_gpu.py tests/test_ctc_gpu.py -q > gpurun_out/r2_test_misc.log 2>&1
echo "misc rc=$?" >> gpurun_out/r2_test_misc.log
tail -15 gpurun_out/r2_test_misc.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_gw8.json 2> gpurun_out/r2_bench_gw8.err
tail -1 gpurun_out/r2_bench_gw8.json | cut -c1-400

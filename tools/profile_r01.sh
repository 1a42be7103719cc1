#!/bin/bash
# ncu evidence for round 1 (run under gpurun, 1 GPU).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of one bench invocation (cold-cache, serialised: compare SHARES)
timeout -s KILL 900 $NCU --metrics gpu__time_duration.sum -s 292 -c 160 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# 2. full-set captures of the top kernels
timeout -s KILL 900 $NCU --set full --import-source on -k regex:lstm_rec_fwd_kernel -s 5 -c 1 -o gpurun_out/prof_rec_fwd_r01 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rec_fwd.log 2>&1
timeout -s KILL 900 $NCU --set full --import-source on -k regex:lstm_rec_bwd_kernel -s 5 -c 1 -o gpurun_out/prof_rec_bwd_r01 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rec_bwd.log 2>&1
timeout -s KILL 900 $NCU --set full --import-source on -k regex:gemm_tc_kernel -s 40 -c 3 -o gpurun_out/prof_gemm_r01 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout -s KILL 900 $NCU --set full --import-source on -k regex:ctc_ -s 9 -c 3 -o gpurun_out/prof_ctc_r01 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ctc.log 2>&1
ls -la gpurun_out/

#!/bin/bash
# ncu evidence for round 2 (run under gpurun, 1 GPU).  Numbers printed under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of one training step (cold-cache, serialised: compare SHARES)
timeout -s KILL 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02f_launches_all.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1
# 2. full-set captures: recurrence kernels (one launch each), CTC kernels at config 4, wide-layer kernels, VGG conv GEMMs
timeout -s KILL 600 $NCU --set full --import-source on -k regex:lstm_rec_fwd_kernel -s 5 -c 1 -f -o gpurun_out/r02f_rec_fwd \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_rec_fwd.log 2>&1
timeout -s KILL 600 $NCU --set full --import-source on -k regex:lstm_rec_bwd_kernel -s 5 -c 1 -f -o gpurun_out/r02f_rec_bwd \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_rec_bwd.log 2>&1
timeout -s KILL 600 $NCU --set full --import-source on -k regex:ctc_ -s 4 -c 4 -f -o gpurun_out/r02f_ctc_cfg4 \
  python tools/prof_ctc_one.py cfg4 > gpurun_out/r02f_ncu_ctc4.log 2>&1
timeout -s KILL 600 $NCU --set full --import-source on -k regex:ctc_ -s 4 -c 4 -f -o gpurun_out/r02f_ctc_cfg2 \
  python tools/prof_ctc_one.py cfg2 > gpurun_out/r02f_ncu_ctc2.log 2>&1
WIDE_T=300 timeout -s KILL 900 $NCU --set full --import-source on -k regex:lstm_wide_ -s 2 -c 2 -f -o gpurun_out/r02f_wide \
  python tools/bench_wide.py > gpurun_out/r02f_ncu_wide.log 2>&1
timeout -s KILL 900 $NCU --set full -k regex:gemm_tc_kernel -s 8 -c 8 -f -o gpurun_out/r02f_vgg_gemm \
  python tools/bench_vgg.py 8 1500 > gpurun_out/r02f_ncu_vgg.log 2>&1
ls -la gpurun_out/ | grep r02f

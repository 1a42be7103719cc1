#!/bin/bash
# ncu --set full of one launch of each recurrence kernel inside a bench step (run under gpurun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout -s KILL 600 $NCU --set full --import-source on -k regex:lstm_rec_fwd_kernel -s 5 -c 1 -f -o gpurun_out/r02f_rec_fwd \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_rec_fwd.log 2>&1
timeout -s KILL 600 $NCU --set full --import-source on -k regex:lstm_rec_bwd_kernel -s 5 -c 1 -f -o gpurun_out/r02f_rec_bwd \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_rec_bwd.log 2>&1
ls -la gpurun_out/r02f_rec_*.ncu-rep

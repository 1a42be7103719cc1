#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in cfg2 cfg4; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ctc_ -s 3 -c 3 -f -o gpurun_out/r2_ctc_$cfg python tools/prof_ctc_one.py $cfg > gpurun_out/r2_ncu17_$cfg.log 2>&1
tail -3 gpurun_out/r2_ncu17_$cfg.log
done
ls -la gpurun_out/*.ncu-rep

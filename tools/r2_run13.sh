#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dp_check.py > gpurun_out/r2_dp13.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dp13.log; grep -v "^\*\*\*\|^$\|OMP_NUM" gpurun_out/r2_dp13.log | head -40

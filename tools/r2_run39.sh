#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_test39.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test39.log; tail -4 gpurun_out/r2_test39.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2_bench39.log 2>&1; tail -1 gpurun_out/r2_bench39.log | cut -c1-400

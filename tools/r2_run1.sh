#!/bin/bash
# round-2 first GPU visit: per-step cycle breakdown of the recurrence kernels + cfg2 parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B2_REC_DBG=1 timeout 300 python tools/bench_rec.py --bwd > gpurun_out/r2_rec_dbg.log 2>&1
timeout 300 python tools/bench_rec.py --bwd > gpurun_out/r2_rec.log 2>&1
timeout 900 python -m pytest tests/test_parity_cfg2_gpu.py -q -s > gpurun_out/r2_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2_parity.log
tail -5 gpurun_out/r2_parity.log

#!/bin/bash
# end-of-round validation under gpurun: full GPU test suite, smoke, bench line, config-3 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/final_tests.log 2>&1
echo "rc=$?" >> gpurun_out/final_tests.log; tail -4 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/final_bench.log 2>&1; grep "^{" gpurun_out/final_bench.log | cut -c1-260
timeout 600 python tools/bench_configs.py cfg3 > gpurun_out/final_cfg3.log 2>&1; tail -4 gpurun_out/final_cfg3.log

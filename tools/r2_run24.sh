#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optim_gpu.py tests/test_model_gpu.py tests/test_towers_gpu.py tests/test_compat_gpu.py tests/test_parity_cfg2_gpu.py -q -x > gpurun_out/r2_test24.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test24.log; tail -5 gpurun_out/r2_test24.log
for v in 1 0; do
B2_BENCH_BUCKETS=$v timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2_bench24_$v.log 2>&1; python - $v <<'PY'
import json, sys
for l in open("gpurun_out/r2_bench24_%s.log" % sys.argv[1]):
    if l.startswith("{"):
        j = json.loads(l); print("bucketed update %s:" % sys.argv[1], round(j["value"]), j["ms_per_step"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["e2e"]["loss"])
PY
done

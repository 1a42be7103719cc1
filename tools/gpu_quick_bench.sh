#!/bin/bash
# targeted GPU tests + two bench lines under gpurun: gpu_quick_bench.sh <pytest args>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest "$@" 2>&1 | tail -4
for i in 1 2; do
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench_$i.log 2>&1; python - $i <<'PY'
import json, sys
for l in open("gpurun_out/quick_bench_%s.log" % sys.argv[1]):
    if l.startswith("{"):
        j = json.loads(l); print("run %s:" % sys.argv[1], round(j["value"]), j["ms_per_step"], j["e2e"]["ms_per_step"], j["gpu_launches"])
PY
done

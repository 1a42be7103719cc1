#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_model_gpu.py -q -x > gpurun_out/r2_test20.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test20.log; tail -8 gpurun_out/r2_test20.log
for v in "8 8" "0 0" "8 0" "16 16" "4 4"; do
set -- $v
B2_FWD_CHUNKS=$1 B2_BWD_CHUNKS=$2 timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench20_$1_$2.log 2>&1; python - $1 $2 <<'PY'
import json, sys
for l in open("gpurun_out/r2_bench20_%s_%s.log" % (sys.argv[1], sys.argv[2])):
    if l.startswith("{"):
        j = json.loads(l); print("fwd chunks %s bwd chunks %s:" % (sys.argv[1], sys.argv[2]), round(j["value"]), j["ms_per_step"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["e2e"]["loss"], j["clocks"])
PY
done

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 1 0 1 0; do
B2_DY_MASK_PASS=$v timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench28_$v.log 2>&1; python - $v <<'PY'
import json, sys
for l in open("gpurun_out/r2_bench28_%s.log" % sys.argv[1]):
    if l.startswith("{"):
        j = json.loads(l); print("dy mask pass %s:" % sys.argv[1], round(j["value"]), j["ms_per_step"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["rooflines"]["blstm_recurrence_bwd"]["ms"])
PY
done
timeout 300 python -m pytest tests/test_lstm_rec_tc_gpu.py -q -x -k "dropout" 2>&1 | tail -2
B2_DY_MASK_PASS=0 timeout 300 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_model_gpu.py -q -x -k "dropout" 2>&1 | tail -2

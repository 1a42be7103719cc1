#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lstm_rec_tc_gpu.py -q -x -k "chunk_chain or forward_backward" > gpurun_out/r2_test19.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test19.log; tail -15 gpurun_out/r2_test19.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench19.log 2>&1; python - <<'PY'
import json
for l in open("gpurun_out/r2_bench19.log"):
    if l.startswith("{"):
        j = json.loads(l); print("chunks on :", j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], j["gpu_launches"])
PY
B2_FWD_CHUNKS=0 timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench19_off.log 2>&1; python - <<'PY'
import json
for l in open("gpurun_out/r2_bench19_off.log"):
    if l.startswith("{"):
        j = json.loads(l); print("chunks off:", j["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], j["gpu_launches"])
PY
tail -3 gpurun_out/r2_bench19.log | cut -c1-300

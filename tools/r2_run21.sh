#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_test21.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test21.log; tail -8 gpurun_out/r2_test21.log
for v in "16 16" "32 32"; do
set -- $v
B2_FWD_CHUNKS=$1 B2_BWD_CHUNKS=$2 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2_bench21_$1_$2.log 2>&1; python - $1 $2 <<'PY'
import json, sys
for l in open("gpurun_out/r2_bench21_%s_%s.log" % (sys.argv[1], sys.argv[2])):
    if l.startswith("{"):
        j = json.loads(l); print("fwd chunks %s bwd chunks %s:" % (sys.argv[1], sys.argv[2]), round(j["value"]), j["ms_per_step"], j["e2e"]["ms_per_step"], j["gpu_launches"], j["e2e"]["loss"], j["rooflines"]["ctc_alpha_beta"]["ms"])
PY
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ctc_ -s 3 -c 3 -f -o gpurun_out/r2_ctc_cfg4_b python tools/prof_ctc_one.py cfg4 > gpurun_out/r2_ncu21.log 2>&1; tail -2 gpurun_out/r2_ncu21.log

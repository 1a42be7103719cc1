#!/bin/bash
# round-2 visit 4: whole step with the deferred side-stream weight-gradient GEMMs on/off; full GPU test suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for ss in 0 1; do
  B2_SIDE_STREAM=$ss timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ss$ss.json 2> gpurun_out/r2_bench_ss$ss.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_ss$ss.json").read().strip().splitlines()[-1])
    print("side_stream=$ss ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "launches", d["gpu_launches"])
    for k,v in d["rooflines"].items(): print("   ", k, round(v.get("ms",0),3), "ms frac", round(v["frac"],4))
except Exception as e:
    print("bench ss=$ss failed", e); print(open("gpurun_out/r2_bench_ss$ss.err").read()[-2000:])
PY
done
B2_SIDE_STREAM=1 B2_BENCH_KEEP_PROB=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ss1_kp1.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_ss1_kp1.json').read().strip().splitlines()[-1]); print('keep_prob=1 side=1 ms/step', d['ms_per_step'])"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_test_full.log 2>&1
echo "full rc=$?" >> gpurun_out/r2_test_full.log
tail -8 gpurun_out/r2_test_full.log

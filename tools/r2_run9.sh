#!/bin/bash
# round-2 visit 9: full GPU suite + default bench (with CPU arm) + reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_test9_full.log 2>&1
echo "full rc=$?" >> gpurun_out/r2_test9_full.log; tail -12 gpurun_out/r2_test9_full.log
grep "^\[parity\|^\[vgg" gpurun_out/r2_test9_full.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke9.log 2>&1; tail -3 gpurun_out/r2_smoke9.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; tail -4 gpurun_out/r2_bench9.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench9.json').read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks']); print('cpu', d['cpu_baseline'])"
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r2_bench9_ref.json 2> gpurun_out/r2_bench9_ref.err; tail -4 gpurun_out/r2_bench9_ref.err; cut -c1-600 gpurun_out/r2_bench9_ref.json

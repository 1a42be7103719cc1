#!/bin/bash
# round-2 visit 7: BPTT with 8 issuer warps + deferred dG stores; TF beam search; full-size parity; new tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lstm_rec_tc_gpu.py tests/test_model_gpu.py tests/test_lstm_gpu.py -x -q > gpurun_out/r2_test7.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test7.log; tail -3 gpurun_out/r2_test7.log
timeout 120 python tools/bench_rec.py --quick > gpurun_out/r2_rec7.log 2>&1; cat gpurun_out/r2_rec7.log
B2_REC_GW=8 timeout 120 python tools/bench_rec.py --quick >> gpurun_out/r2_rec7.log 2>&1; tail -2 gpurun_out/r2_rec7.log
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_compat_gpu.py -q > gpurun_out/r2_test7b.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test7b.log; tail -6 gpurun_out/r2_test7b.log
timeout 1500 python -m pytest tests/test_parity_fullsize_gpu.py -q -s > gpurun_out/r2_test7c.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test7c.log; grep "^\[parity" gpurun_out/r2_test7c.log | cut -c1-700; tail -6 gpurun_out/r2_test7c.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench7.json').read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])
for k,v in d['rooflines'].items(): print('   ', k, round(v.get('ms',0),3), 'ms frac', round(v['frac'],4))"

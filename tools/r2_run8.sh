#!/bin/bash
# round-2 visit 8: VGG on tensor cores (tests + cfg4 front-end timing), cfg5 parity, bench sanity, ncu launch list + captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vgg_gpu.py tests/test_gemm_gpu.py -q > gpurun_out/r2_test8.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test8.log; tail -12 gpurun_out/r2_test8.log
timeout 300 python tools/bench_vgg.py 32 1500 > gpurun_out/r2_vgg8.log 2>&1; cat gpurun_out/r2_vgg8.log
timeout 900 python -m pytest tests/test_parity_fullsize_gpu.py -q -s -k cfg5 > gpurun_out/r2_test8c.log 2>&1
echo "rc=$?" >> gpurun_out/r2_test8c.log; grep "^\[parity" gpurun_out/r2_test8c.log | cut -c1-400; tail -3 gpurun_out/r2_test8c.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench8.json').read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks'], 'launches', d['gpu_launches'])"
# launch list of one training step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench8.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r02_launches_raw.csv
# one full capture each of the two recurrence kernels (layer-sized launch)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_bwd -s 2 -c 1 -o gpurun_out/prof_rec_bwd_r02 python tools/bench_rec.py --quick > gpurun_out/ncu_rec_bwd8.log 2>&1
echo "ncu bwd rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_fwd -s 2 -c 1 -o gpurun_out/prof_rec_fwd_r02 python tools/bench_rec.py --quick > gpurun_out/ncu_rec_fwd8.log 2>&1
echo "ncu fwd rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4

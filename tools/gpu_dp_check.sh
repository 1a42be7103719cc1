#!/bin/bash
# N-GPU validation under gpurun --gpus N: exchange modes agree, bench line at N
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 tools/dp_check.py > gpurun_out/dp_check_n$N.log 2>&1; echo "rc=$?"; grep -v "^\*\*\*\|^$\|OMP_NUM" gpurun_out/dp_check_n$N.log | tail -6
timeout 600 $TR --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "rc=$?"; grep "^{" gpurun_out/bench_n$N.log | python -c "
import json,sys
for l in sys.stdin:
    j=json.loads(l); print(j['n_gpus'], round(j['value']), j['ms_per_step'], j['e2e']['ms_per_step'], j['config']['gradient_exchange'], j['scaling'])
"

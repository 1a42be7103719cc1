#!/bin/bash
# ncu --set full captures of the kernels outside the headline step (run under gpurun, 1 GPU)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout -s KILL 300 $NCU -k regex:attention_ -s 4 -c 5 -o gpurun_out/prof_attn_r01 python tools/prof_extra.py attn > gpurun_out/ncu_attn.log 2>&1
timeout -s KILL 300 $NCU -k regex:gemm_skinny -s 2 -c 2 -o gpurun_out/prof_gemv_r01 python tools/prof_extra.py gemv > gpurun_out/ncu_gemv.log 2>&1
timeout -s KILL 300 $NCU -k regex:stack_splice -s 1 -c 1 -o gpurun_out/prof_inputs_r01 python tools/prof_extra.py inputs > gpurun_out/ncu_inputs.log 2>&1
timeout -s KILL 300 $NCU -k regex:sequence_loss -s 1 -c 1 -o gpurun_out/prof_seqloss_r01 python tools/prof_extra.py seqloss > gpurun_out/ncu_seqloss.log 2>&1
ls -la gpurun_out/*.ncu-rep

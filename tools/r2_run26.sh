#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/probe_rec_linear.py > gpurun_out/r2_probe26.log 2>&1; cat gpurun_out/r2_probe26.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/prof_cfg4_host.py > gpurun_out/r2_host33.log 2>&1; grep -v "^$" gpurun_out/r2_host33.log | head -40

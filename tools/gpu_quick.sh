#!/bin/bash
# run a subset of the GPU tests under gpurun: gpu_quick.sh <pytest args>
cd "$(dirname "$0")/.."
timeout 900 python -m pytest "$@" 2>&1 | tail -12

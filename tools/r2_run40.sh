#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_vgg_gpu.py -q -x -s -k "wide_blstm" 2>&1 | tail -15

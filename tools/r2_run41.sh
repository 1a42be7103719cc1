#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_lstm_encoder_gpu.py -q -x 2>&1 | tail -15

#!/bin/bash
# the CLI suite on the real engine with the final host binary (pipes are spooled by default now)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_cli_gpu.py -q -m gpu --timeout 120 --durations=8 2>&1 | tail -16 ) 2>&1 | tail -20

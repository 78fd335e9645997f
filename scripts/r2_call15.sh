#!/bin/bash
cd "$(dirname "$0")/.."
timeout 500 python -m pytest tests/test_cli_gpu.py -x -q -m gpu --timeout 150 -k "sharded" 2>&1 | tail -25

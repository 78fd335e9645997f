#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_fast_gpu.py tests/test_seq_gpu.py tests/test_unordered_gpu.py tests/test_sharded2_gpu.py tests/test_sharded_gpu.py -x -q -m gpu --timeout 100 2>&1 | tail -8
timeout 200 scripts/r2_k1_batch.sh

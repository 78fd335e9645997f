#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_sharded_unordered_gpu.py tests/test_unordered_gpu.py tests/test_sharded_seq_gpu.py -x -q -m gpu --timeout 200 2>&1 | tail -25
echo "== tail-hamming trace"
FQD_TRACE=1 FQD_TRACE_SORT=1 timeout 200 python bench_seq.py --mode tail-hamming --pairs 50000000 --steps 1 2>&1 | grep -E "fqd trace" | tail -12

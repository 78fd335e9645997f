#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_fast_gpu.py tests/test_seq_gpu.py tests/test_unordered_gpu.py -x -q -m gpu --timeout 60 --durations=8 2>&1 | tail -25
timeout 400 scripts/r2_k1_batch.sh
echo "== ncu K1"
export FQD_BENCH_READS=6000000 FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_CPU=1 FQD_BENCH_SKIP_MODES=1
timeout 120 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_parse_pack -s 3 -c 1 -f -o gpurun_out/prof_k1_r2 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_k1.log 2>&1
tail -3 gpurun_out/ncu_k1.log

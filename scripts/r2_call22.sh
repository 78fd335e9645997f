#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fast_gpu.py tests/test_unordered_gpu.py -x -q -m gpu --timeout 100 2>&1 | tail -3
export FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_CPU=1 FQD_BENCH_SKIP_MODES=1
for mode in two one; do
  if [ $mode = one ]; then export FQD_K2_ONEPASS=1; fi
  out=$(timeout 200 python bench.py --steps 10 --warmup 3 2>gpurun_out/err_k2_$mode.log)
  echo "K2 $mode-pass: $(echo "$out" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('step_ms', round(d['ms_per_step'],3), 'value', round(d['value']/1e9,3), 'k1_share', round(r['kernel_share_of_step'],3), 'k2_share', round(r['insert_share_of_step'],3), 'k2_ms_per_job', round(r['insert_share_of_step']*d['ms_per_step'],2), 'dups', d['duplicates_removed'])")"
done
unset FQD_K2_ONEPASS
echo "== seq trace (tight, 50M pairs)"
FQD_TRACE=1 timeout 200 python bench_seq.py --mode tight --pairs 50000000 --steps 1 2>&1 | grep -E "fqd trace|value" | tail -40 | cut -c1-200

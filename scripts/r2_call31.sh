#!/bin/bash
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests/test_seq_gpu.py tests/test_unordered_gpu.py -x -q -m gpu --timeout 100 2>&1 | tail -3
for m in tight loose tail-hamming unordered; do
  timeout 200 python bench_seq.py --mode $m --pairs 50000000 --steps 2 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['mode'], 'ms', round(d['ms_per_step'],2), 'Gpairs/s', round(d['value']/1e9,3), 'dups', d['duplicates_removed'], 'out', d['pairs_out'])"
done
echo "== tail-hamming trace"
FQD_TRACE=1 FQD_TRACE_SORT=1 timeout 200 python bench_seq.py --mode tail-hamming --pairs 50000000 --steps 1 2>&1 | grep -E "fqd trace" | tail -13

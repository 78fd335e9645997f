#!/bin/bash
cd "$(dirname "$0")/.."
for m in loose tail-hamming unordered; do
  echo "== $m"
  FQD_TRACE=1 timeout 200 python bench_seq.py --mode $m --pairs 50000000 --steps 1 2>&1 | grep -E "fqd trace" | tail -14
done

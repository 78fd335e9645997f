#!/bin/bash
# usage (under gpurun): scripts/r2_k1_batch.sh  - K1/K2 experiment builds, one bench line each (20 M reads, no e2e / cpu legs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export FQD_BENCH_READS=${FQD_BENCH_READS:-20000000} FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_CPU=1
V=fastq-dupaway_b200/csrc/variants
for lib in fastq-dupaway_b200/csrc/libfqd_cuda.so $V/*.so; do
  name=$(basename $lib .so)
  case $name in *timeline*) continue;; esac
  out=$(FQD_LIB=$lib timeout 120 python bench.py --steps 5 --warmup 3 2>gpurun_out/err_$name.log)
  echo "$name rc $? $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('k1_ms', round(r['avg_launch_ms'],4), 'frac', round(r['frac'],4), 'k1_share', round(r['kernel_share_of_step'],3), 'k2_share', round(r['insert_share_of_step'],3), 'step_ms', round(d['ms_per_step'],3), 'dups', d['duplicates_removed'])
except Exception as e: print('no json', e)")" | tee -a gpurun_out/k1_batch.txt
done
for lib in $V/*timeline*.so; do
  [ -f "$lib" ] || continue
  echo "== $(basename $lib)" | tee -a gpurun_out/k1_batch.txt
  FQD_LIB=$lib timeout 120 python scripts/k1_timeline.py 2>&1 | tee -a gpurun_out/k1_batch.txt
done

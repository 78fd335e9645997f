#!/bin/bash
# launch lists of two whole-input jobs on the final library (bench_seq.py, 5 M pairs; the same command has just exited 0 without ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in tight unordered; do
  timeout 200 python bench_seq.py --mode $m --pairs 5000000 --steps 1 > gpurun_out/seq_$m.json 2>&1 && \
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_seq_${m}_5Mpairs.csv \
      python bench_seq.py --mode $m --pairs 5000000 --steps 1 > gpurun_out/ncu_$m.log 2>&1
  echo "$m list rc=$? lines=$(wc -l < gpurun_out/r02_launches_seq_${m}_5Mpairs.csv)"
done

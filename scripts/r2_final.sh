#!/bin/bash
# round-2 closing run on one B200: the whole GPU suite, smoke(), then the launch lists (ncu, one metric, no replay) of
# the bench command and of two whole-input jobs - each only after the same command has exited 0 without ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -6 ) 2>&1 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
export FQD_BENCH_READS=20000000 FQD_BENCH_SKIP_MODES=1 FQD_BENCH_SKIP_CPU=1 FQD_BENCH_SKIP_E2E=1
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fast.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_fast.log 2>&1
echo "fast list rc=$? lines=$(wc -l < gpurun_out/r02_launches_fast.csv)"
for m in tail-hamming unordered; do
  timeout 200 python bench_seq.py --mode $m --pairs 5000000 --steps 1 > gpurun_out/seq_$m.json 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/r02_launches_seq_${m}_5Mpairs.csv \
      python bench_seq.py --mode $m --pairs 5000000 --steps 1 > gpurun_out/ncu_$m.log 2>&1
  echo "$m list rc=$? lines=$(wc -l < gpurun_out/r02_launches_seq_${m}_5Mpairs.csv)"
done

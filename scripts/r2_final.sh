#!/bin/bash
# round-2 closing run on one B200: launch lists (ncu, one metric, no replay) of the bench command and of two whole-input
# jobs - each only after the same command has exited 0 without ncu -, smoke(), the GPU suite (the parts this session's
# earlier calls have not already run on the final library), then the default bench.py line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export FQD_BENCH_READS=20000000 FQD_BENCH_SKIP_MODES=1 FQD_BENCH_SKIP_CPU=1 FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_FILES=1
timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_fast.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_fast.log 2>&1
echo "fast list rc=$? lines=$(wc -l < gpurun_out/r02_launches_fast.csv)"
for m in tight unordered; do
  timeout 200 python bench_seq.py --mode $m --pairs 5000000 --steps 1 > gpurun_out/seq_$m.json 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/r02_launches_seq_${m}_5Mpairs.csv \
      python bench_seq.py --mode $m --pairs 5000000 --steps 1 > gpurun_out/ncu_$m.log 2>&1
  echo "$m list rc=$? lines=$(wc -l < gpurun_out/r02_launches_seq_${m}_5Mpairs.csv)"
done
unset FQD_BENCH_READS FQD_BENCH_SKIP_MODES FQD_BENCH_SKIP_CPU FQD_BENCH_SKIP_E2E FQD_BENCH_SKIP_FILES
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python -m pytest tests -q -m gpu --timeout 300 -k "not discard" --ignore=tests/test_cli_gpu.py > gpurun_out/final_gpu_tests.log 2>&1 ) 2>&1 | grep real
tail -4 gpurun_out/final_gpu_tests.log
( time timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real
tail -2 gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json

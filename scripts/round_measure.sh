timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full_v8.json 2> gpurun_out/bench_full_v8.err
tail -2 gpurun_out/bench_full_v8.err
FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_CPU=1 FQD_BENCH_READS=20000000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fast_v8.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_fast_v8.log 2>&1
FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_CPU=1 FQD_BENCH_READS=6000000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_insert -s 3 -c 1 -f -o gpurun_out/prof_insert_v8 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_insert_v8.log 2>&1
timeout 900 python bench_seq.py --pairs 50000000 --steps 2 --cpu > gpurun_out/bench_seq_50M.json 2> gpurun_out/bench_seq_50M.err
tail -2 gpurun_out/bench_seq_50M.err

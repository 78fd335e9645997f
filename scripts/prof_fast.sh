export FQD_BENCH_READS=6000000 FQD_BENCH_SKIP_E2E=1 FQD_BENCH_SKIP_CPU=1
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_parse_pack -s 3 -c 1 -o gpurun_out/prof_parse python bench.py --steps 2 --warmup 3 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_insert -s 3 -c 1 -o gpurun_out/prof_insert python bench.py --steps 2 --warmup 3 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/plain.log gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log

# CLI GPU tests, files-in/files-out bench of the drop-in binary vs the reference binary, and the default bench line.
nproc; lscpu | grep -E "Model name|Socket|Thread|Core" | head -5
timeout 600 python -m pytest tests/test_cli_gpu.py -x -q -m gpu > gpurun_out/cli_tests.log 2>&1; tail -3 gpurun_out/cli_tests.log
timeout 500 python bench_cli.py --reads ${CLI_READS:-20000000} --ref-reads 2000000 > gpurun_out/bench_cli.json 2> gpurun_out/bench_cli.err; tail -3 gpurun_out/bench_cli.err; cut -c1-330 gpurun_out/bench_cli.json
FQD_IO_THREADS=1 timeout 300 python bench_cli.py --reads ${CLI_READS:-20000000} --ref-reads 100000 --formats plain,bgzf --repeats 1 > gpurun_out/bench_cli_1thread.json 2> gpurun_out/bench_cli_1thread.err; cut -c1-330 gpurun_out/bench_cli_1thread.json
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; tail -2 gpurun_out/bench_r1_final2.err; cut -c1-200 gpurun_out/bench_r1_final2.json

# CLI GPU tests, then files-in/files-out bench of the drop-in binary vs the reference binary.
nproc; lscpu | grep -E "Model name|Socket|Thread|Core" | head -5
timeout 400 python -m pytest tests/test_cli_gpu.py -x -q -m gpu > gpurun_out/cli_tests.log 2>&1; tail -3 gpurun_out/cli_tests.log
timeout 400 python bench_cli.py --reads ${CLI_READS:-20000000} --ref-reads 2000000 --formats ${CLI_FORMATS:-plain,gz1,bgzf} > gpurun_out/bench_cli.json 2> gpurun_out/bench_cli.err; tail -3 gpurun_out/bench_cli.err; cut -c1-360 gpurun_out/bench_cli.json

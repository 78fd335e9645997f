N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_fast_N$N.json 2> gpurun_out/bench_fast_N$N.err
grep -v "^\*\|OMP_NUM" gpurun_out/bench_fast_N$N.err | tail -3; cut -c1-200 gpurun_out/bench_fast_N$N.json

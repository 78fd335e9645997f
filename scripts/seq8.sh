N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench_seq.py --pairs 20000000 --steps 2 > gpurun_out/bench_seq_N$N.json 2> gpurun_out/bench_seq_N$N.err
grep -v "^\*\|OMP_NUM" gpurun_out/bench_seq_N$N.err | tail -3; cut -c1-230 gpurun_out/bench_seq_N$N.json

# 2-GPU check of both multi-GPU arms (run under gpurun --gpus 2)
export FQD_BENCH_READS=${FQD_BENCH_READS:-40000000}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_fast_N2.json 2> gpurun_out/bench_fast_N2.err
tail -3 gpurun_out/bench_fast_N2.err; cut -c1-300 gpurun_out/bench_fast_N2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_seq.py --pairs ${PAIRS:-20000000} --steps 2 > gpurun_out/bench_seq_N2.json 2> gpurun_out/bench_seq_N2.err
tail -5 gpurun_out/bench_seq_N2.err; cut -c1-400 gpurun_out/bench_seq_N2.json

# launch list of the sharded --fast path with one rank (kernel costs of the exchange stages; run under gpurun)
export FQD_BENCH_READS=20000000 FQD_BENCH_FORCE_SHARDED=1 RANK=0 LOCAL_RANK=0 WORLD_SIZE=1 MASTER_ADDR=127.0.0.1 MASTER_PORT=29533
timeout 300 python bench.py --gpus 1 --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_sharded1.csv python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/ncu_sharded1.log 2>&1
tail -2 gpurun_out/ncu_sharded1.log | cut -c1-200

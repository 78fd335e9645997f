export FQD_BENCH_READS=40000000
for ch in 4000000 10000000 20000000; do
FQD_BENCH_CHUNK_READS=$ch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('chunk $ch', 'ms', round(d['ms_per_step'],2), 'Greads/s', round(d['value']/1e9,3), 'k1', round(d['roofline']['kernel_share_of_step'],3), 'ins', round(d['roofline']['insert_share_of_step'],3))"
done

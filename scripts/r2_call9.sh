#!/bin/bash
# usage (gpurun --gpus 2): sharded --fast at N=2, full size, then the GPU tests of the sharded paths
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_full_n$N.json 2> gpurun_out/bench_full_n$N.err ) 2>&1 | tail -4
tail -15 gpurun_out/bench_full_n$N.err
python - $N <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_full_n{n}.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "k1_ms", d["roofline"]["avg_launch_ms"], "shares", d["roofline"]["kernel_share_of_step"], d["roofline"]["insert_share_of_step"], "verify", d["verify"], "launches", d["gpu_launches"])
    for m, v in (d.get("modes") or {}).items():
        print(m, {k: v.get(k) for k in ("value", "ms_per_step", "error", "leg_wall_s", "verify", "imbalance")}, "frac", (v.get("roofline") or {}).get("frac"))
except Exception as e:
    print("no json", e)
P

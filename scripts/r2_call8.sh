#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 scripts/r2_k1_batch.sh
echo "== full bench N=1"
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err ) 2>&1 | tail -4; tail -5 gpurun_out/bench_full_n1.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench_full_n1.json").read())
    print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "k1_ms", d["roofline"]["avg_launch_ms"], "shares", d["roofline"]["kernel_share_of_step"], d["roofline"]["insert_share_of_step"], "whole", d["roofline"]["whole_path"]["frac"], "e2e", d["e2e"].get("value"), "parity", d["parity"].get("ok"), "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
    for m, v in (d.get("modes") or {}).items():
        print(m, {k: v.get(k) for k in ("value", "ms_per_step", "error", "leg_wall_s")}, "frac", (v.get("roofline") or {}).get("frac"),
              "e2e", (v.get("e2e") or {}).get("value"), (v.get("e2e") or {}).get("error"), "parity", (v.get("parity") or {}).get("ok"), "cpu", (v.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("no json", e)
P

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/bench_full_n1.json 2> gpurun_out/bench_full_n1.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_full_n1.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_full_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "parity", d.get("parity"))
for m, v in (d.get("modes") or {}).items():
    print(m, {k: v.get(k) for k in ("value", "ms_per_step", "error")}, "frac", (v.get("roofline") or {}).get("frac"), "e2e", (v.get("e2e") or {}).get("value"), "cpu", (v.get("cpu_baseline") or {}).get("value"), "parity", v.get("parity"))
P

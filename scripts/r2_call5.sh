#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_fast_gpu.py tests/test_seq_gpu.py tests/test_unordered_gpu.py -x -q -m gpu --timeout 120 2>&1 | tail -15
timeout 400 scripts/r2_k1_batch.sh
echo "== sharded2 on one GPU"
timeout 400 python -m pytest tests/test_sharded2_gpu.py -x -q -m gpu --timeout 150 2>&1 | tail -15
echo "== small bench with modes"
FQD_BENCH_READS=12000000 FQD_BENCH_PAIRS=4000000 FQD_BENCH_E2E_PAIRS=2000000 FQD_BENCH_PARITY_PAIRS=300000 FQD_BENCH_E2E_READS=6000000 \
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc $?"; tail -5 gpurun_out/bench_small.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench_small.json").read())
    print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"].get("value"), "parity", d["parity"])
    for m, v in (d.get("modes") or {}).items():
        print(m, {k: v.get(k) for k in ("value", "ms_per_step", "error", "leg_wall_s")}, "frac", (v.get("roofline") or {}).get("frac"),
              "e2e", (v.get("e2e") or {}).get("value"), (v.get("e2e") or {}).get("error"), "parity", (v.get("parity") or {}).get("ok"), "cpu", (v.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("no json", e)
P

"""bench.py's reference arm runs on the CPU, so its side of the measurement contract can be checked here: one JSON
line with the agreed keys, the same metric / unit / config as the GPU arm, rank > 0 exits without work."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def run_bench(extra_env=None, *args):
    env = dict(os.environ, FQD_REF_BUDGET_S="4")
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", *args],
                          capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    p = run_bench()
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "dedup reads/sec" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and d["n_gpus"] == 1 and d["steps"] == 1
    assert "BASELINE configs[1]" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 10_000                      # a single core manages a few hundred thousand reads per second


def test_reference_arm_other_ranks_do_nothing():
    p = run_bench({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert p.returncode == 0 and p.stdout.strip() == ""

#!/bin/bash
# diagnostic: device-side phase trace ([fqd trace]) of a 60 M-pair tight job through the binary (discarded-input path)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 70 python - <<'P' 2>&1 | tail -60
import importlib, os, subprocess, sys, tempfile, shutil, time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
sys.path.insert(0, os.getcwd())
bc = importlib.import_module("bench_cli")
tmp = Path(tempfile.mkdtemp(prefix="fqd_tr_", dir="/dev/shm"))
try:
    n = 60_000_000
    files = [tmp / "a.fq", tmp / "b.fq"]
    with ThreadPoolExecutor(2) as ex:
        list(ex.map(lambda a: bc.synth_file(a[0], n, a[1]), zip(files, (1, 2))))
    t0 = time.perf_counter()
    r = subprocess.run([str(bc.EXE), "-i", files[0], "-u", files[1], "-o", "/dev/null", "-p", "/dev/zero", "--compare-seq", "tight", "-v"],
                       capture_output=True, text=True, env=dict(os.environ, FQD_TRACE="1", FQD_TRACE_SORT="1"))
    print("wall", round(time.perf_counter() - t0, 2), "rc", r.returncode, r.stdout.strip())
    lines = [l for l in r.stderr.splitlines() if l.startswith("[fqd trace]") or l.startswith("[host-trace]")]
    # the per-segment parse marks repeat: keep the last 45 lines (finish stages) and the host marks
    for l in [l for l in lines if l.startswith("[host-trace]")] + lines[-45:]:
        print(l)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
P

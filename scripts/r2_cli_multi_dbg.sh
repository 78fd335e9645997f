#!/bin/bash
cd "$(dirname "$0")/.."
python - <<'P'
import importlib, sys, subprocess, os
sys.path.insert(0, ".")
bc = importlib.import_module("bench_cli")
from pathlib import Path
tmp = Path("/dev/shm/fqd_dbg"); tmp.mkdir(exist_ok=True)
for m in (1, 2):
    bc.synth_file(tmp / f"r{m}.fq", 2_000_000, m)
for devs in ("0,1",):
    r = subprocess.run([str(bc.EXE), "-i", tmp / "r1.fq", "-u", tmp / "r2.fq", "-o", tmp / "o1.fq", "-p", tmp / "o2.fq", "--fast", "-v"],
                       capture_output=True, text=True, env=dict(os.environ, FQD_DEVICES=devs, FQD_TRACE="1", FQD_BACKTRACE="1"))
    print("devices", devs, "rc", r.returncode, "stdout", r.stdout.strip(), "stderr tail:", r.stderr[-2500:].replace("\n", " | "), flush=True)
r = subprocess.run([str(bc.EXE), "-i", tmp / "r1.fq", "-o", tmp / "o1.fq", "--fast", "-v"], capture_output=True, text=True, env=dict(os.environ, FQD_DEVICE="1"))
print("single on device 1: rc", r.returncode, r.stdout.strip(), r.stderr[-300:])
import shutil; shutil.rmtree(tmp)
P

#!/usr/bin/env python
"""The drop-in binary with FQD_DEVICES (duplicate set sharded over several GPUs) against the same binary on one GPU and
against the reference binary on a prefix: paired-end FASTQ files on tmpfs, `--fast -v`.  One JSON line.

    python scripts/r2_cli_multi.py [--pairs 20000000] [--ref-pairs 1500000] [--devices 0,1]
"""
import argparse
import hashlib
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
bc = importlib.import_module("bench_cli")


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for b in iter(lambda: f.read(1 << 24), b""):
            h.update(b)
    return h.hexdigest()


def run(cmd, env=None, cwd=None):
    t0 = time.perf_counter()
    r = subprocess.run(list(map(str, cmd)), capture_output=True, text=True, env=dict(os.environ, **(env or {})), cwd=cwd)
    return time.perf_counter() - t0, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=20_000_000)
    ap.add_argument("--ref-pairs", type=int, default=1_500_000)
    ap.add_argument("--devices", default="0,1")
    a = ap.parse_args()
    oracle = importlib.import_module("oracle")
    tmp = Path(tempfile.mkdtemp(prefix="fqd_cli_", dir="/dev/shm"))
    try:
        for m in (1, 2):
            bc.synth_file(tmp / f"r{m}.fq", a.pairs, m)
        out = {"pairs": a.pairs, "input_bytes": 2 * a.pairs * bc.REC, "devices": a.devices}
        io = lambda tag: ["-i", tmp / "r1.fq", "-u", tmp / "r2.fq", "-o", tmp / f"{tag}1.fq", "-p", tmp / f"{tag}2.fq", "--fast", "-v"]
        t1, r1 = run([bc.EXE, *io("one")])
        assert r1.returncode == 0, r1.stderr
        tn, rn = run([bc.EXE, *io("many")], env={"FQD_DEVICES": a.devices})
        if rn.returncode != 0:
            print("SHARDED FAILED:", rn.stderr[-2000:], flush=True)
        assert rn.returncode == 0
        same = all(sha(tmp / f"one{m}.fq") == sha(tmp / f"many{m}.fq") for m in (1, 2)) and r1.stdout == rn.stdout
        out.update({"one_gpu_s": t1, "one_gpu_pairs_per_s": a.pairs / t1, "sharded_s": tn, "sharded_pairs_per_s": a.pairs / tn,
                    "summary_line": rn.stdout.strip(), "outputs_identical_one_vs_sharded": same})
        # reference binary on a prefix, and the sharded binary on the same prefix: bytes must be identical
        nb = a.ref_pairs * bc.REC
        for m in (1, 2):
            with open(tmp / f"r{m}.fq", "rb") as f, open(tmp / f"p{m}.fq", "wb") as g:
                g.write(f.read(nb))
        tr, rr = run([oracle.REF_BIN, "-i", "p1.fq", "-u", "p2.fq", "-o", "ref1.fq", "-p", "ref2.fq", "--fast", "-v"], cwd=tmp)
        assert rr.returncode == 0, rr.stderr
        tp, rp = run([bc.EXE, "-i", tmp / "p1.fq", "-u", tmp / "p2.fq", "-o", tmp / "sp1.fq", "-p", tmp / "sp2.fq", "--fast", "-v"], env={"FQD_DEVICES": a.devices})
        assert rp.returncode == 0, rp.stderr
        ident = all(sha(tmp / f"ref{m}.fq") == sha(tmp / f"sp{m}.fq") for m in (1, 2)) and rr.stdout == rp.stdout
        out.update({"reference_prefix_pairs": a.ref_pairs, "reference_s": tr, "reference_pairs_per_s": a.ref_pairs / tr,
                    "sharded_on_prefix_s": tp, "outputs_byte_identical_to_reference_on_prefix": ident})
        print(json.dumps(out), flush=True)
        assert same and ident
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()

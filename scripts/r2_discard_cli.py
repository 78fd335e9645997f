"""Whole-input modes through the drop-in binary when the raw input does NOT stay on the device (FQD_WHOLE_INPUT=discard,
host/replay.hpp): 2 x 150 bp pairs of bench.py's synthetic stream on tmpfs, --compare-seq <mode> -v.

  1. --pairs P (default 40 M): the resident path and the discarded-input path on the same files - times, phases, device memory
     high-water mark, outputs compared byte by byte; the first --ref-pairs of the same files through the stable-sort build
     of the reference (oracle/_ref) and through the discarded-input path, outputs compared.
  2. --big-pairs B (default: what the box's RAM allows, at most 200 M = BASELINE configs[2]): the policy left to the binary
     (it must pick the discarded-input path by itself when the input does not fit), one run, -v line and phases reported.
One JSON line per run.
"""
from __future__ import annotations

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
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
bench_cli = importlib.import_module("bench_cli")
EXE, REC = bench_cli.EXE, bench_cli.REC


def digest(path: Path) -> str:
    h = hashlib.blake2b(digest_size=16)
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 26)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def run(ins, outs, mode, policy):
    env = dict(os.environ, FQD_TRACE="1")
    env.pop("FQD_WHOLE_INPUT", None)
    if policy:
        env["FQD_WHOLE_INPUT"] = policy
    t0 = time.perf_counter()
    res = subprocess.run([str(EXE), "-i", str(ins[0]), "-u", str(ins[1]), "-o", str(outs[0]), "-p", str(outs[1]), "--compare-seq", mode, "-v"],
                         capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    marks, hw = {}, None
    for line in res.stderr.splitlines():
        if line.startswith("[host-trace] device memory high-water mark"):
            hw = float(line.split("mark", 1)[1].split("GiB", 1)[0])
        elif line.startswith("[host-trace]"):
            ms, label = line[len("[host-trace]"):].split("ms", 1)
            marks[label.strip()] = float(ms)
    assert res.returncode == 0, res.stderr[-3000:]
    def at(prefix):
        return next((v for k, v in marks.items() if k.startswith(prefix)), None)
    ph = {}
    if at("engine created") is not None and at("outputs closed") is not None:
        ph = {"startup_s": round(at("engine created") / 1e3, 2), "ingest_s": round((at("input on the device") - at("engine created")) / 1e3, 2),
              "sort_scan_s": round((at("sorted / joined / scanned") - at("input on the device")) / 1e3, 2),
              "output_s": round((at("outputs closed") - at("sorted / joined / scanned")) / 1e3, 2)}
    discarded = any("raw input not kept" in k for k in marks)
    return dt, res.stdout.strip(), ph, hw, discarded


def make_inputs(tmp: Path, pairs: int, tag: str):
    files = [tmp / f"{tag}_{m}.fq" for m in (1, 2)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=2) as ex:
        list(ex.map(lambda a: bench_cli.synth_file(a[0], pairs, a[1]), zip(files, (1, 2))))
    return files, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=40_000_000)
    ap.add_argument("--ref-pairs", type=int, default=1_000_000)
    ap.add_argument("--big-pairs", type=int, default=-1, help="-1: from MemAvailable (at most 200 M); 0: skip")
    ap.add_argument("--mode", default="tight")
    ap.add_argument("--big-out", default="tmpfs", choices=["tmpfs", "null"],
                    help="null: the large job writes to /dev/null (input alone fills the box's RAM); records are still gathered")
    args = ap.parse_args()
    oracle = importlib.import_module("oracle")
    tmp = Path(tempfile.mkdtemp(prefix="fqd_discard_", dir="/dev/shm"))
    try:
        if args.pairs > 0:
            ins, gen_s = make_inputs(tmp, args.pairs, "in")
            out = {p: [tmp / f"out_{p}_{m}.fq" for m in (1, 2)] for p in ("resident", "discard")}
            dig = {}
            for policy in ("resident", "discard"):
                dt, so, ph, hw, disc = run(ins, out[policy], args.mode, policy)
                assert disc == (policy == "discard")
                with ThreadPoolExecutor(max_workers=2) as ex:
                    dig[policy] = list(ex.map(digest, out[policy]))
                print(json.dumps({"what": "same files, both paths", "policy": policy, "mode": args.mode, "pairs": args.pairs,
                                  "input_bytes": sum(f.stat().st_size for f in ins), "output_bytes": sum(f.stat().st_size for f in out[policy]),
                                  "seconds": round(dt, 2), "pairs_per_s": round(args.pairs / dt), **ph, "device_GiB_high_water": hw,
                                  "stdout": so, "generate_s": round(gen_s, 1)}), flush=True)
                for f in out[policy]:
                    f.unlink()
            print(json.dumps({"what": "outputs of the two paths", "byte_identical": dig["resident"] == dig["discard"], "blake2b": dig["discard"]}), flush=True)
            # the reference on a prefix against the discarded-input path
            n = min(args.ref_pairs, args.pairs)
            if n > 0 and oracle.ref_available(stable=True):
                pre = [tmp / f"pre_{m}.fq" for m in (1, 2)]
                for src, dst in zip(ins, pre):
                    with open(src, "rb") as f, open(dst, "wb") as g:
                        g.write(f.read(n * REC))
                o_ref = [tmp / f"ref_{m}.fq" for m in (1, 2)]
                o_us = [tmp / f"us_{m}.fq" for m in (1, 2)]
                t0 = time.perf_counter()
                r = subprocess.run([str(oracle.REF_STABLE_BIN), "-i", str(pre[0]), "-u", str(pre[1]), "-o", str(o_ref[0]), "-p", str(o_ref[1]),
                                    "--compare-seq", args.mode, "-v", "-m", "10240"], capture_output=True, text=True, cwd=tmp)
                dt_r = time.perf_counter() - t0
                dt, so, _, _, disc = run(pre, o_us, args.mode, "discard")
                print(json.dumps({"what": "prefix vs oracle/_ref/fastq-dupaway-stable", "pairs": n, "reference_seconds": round(dt_r, 2), "ours_seconds": round(dt, 2),
                                  "outputs_byte_identical": all(a.read_bytes() == b.read_bytes() for a, b in zip(o_us, o_ref)),
                                  "verbose_lines_identical": so == r.stdout.strip(), "discarded_input_path": disc}), flush=True)
            for f in tmp.iterdir():
                f.unlink()
        big = args.big_pairs
        if big < 0:
            avail = 0
            for line in open("/proc/meminfo"):
                if line.startswith("MemAvailable:"):
                    avail = int(line.split()[1]) * 1024
            # input + output on tmpfs (the mapping shares the input's pages): 644 + ~451 bytes per pair, 25 % head room
            for cg in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):      # a container's own limit
                try:
                    v = open(cg).read().strip()
                    if v.isdigit():
                        used = 0
                        try:
                            used = int(open("/sys/fs/cgroup/memory.current").read())
                        except OSError:
                            pass
                        avail = min(avail, int(v) - used)
                except OSError:
                    pass
            shm = shutil.disk_usage("/dev/shm").free
            per_pair = 2 * REC * (1.0 if args.big_out == "null" else 1.7)
            big = min(200_000_000, int(min(avail * (0.8 if args.big_out == "null" else 0.75), shm * 0.9) / per_pair) // 1_000_000 * 1_000_000)
            print(json.dumps({"what": "box", "MemAvailable_GiB": round(avail / 2**30, 1), "dev_shm_free_GiB": round(shm / 2**30, 1), "big_pairs": big,
                              "cores": os.cpu_count()}), flush=True)
        if big > 0:
            ins, gen_s = make_inputs(tmp, big, "big")
            outs = [tmp / f"bigout_{m}.fq" for m in (1, 2)] if args.big_out == "tmpfs" else [Path("/dev/null"), Path("/dev/zero")]
            dt, so, ph, hw, disc = run(ins, outs, args.mode, None)
            print(json.dumps({"what": "large job, policy chosen by the binary", "mode": args.mode, "pairs": big, "input_bytes": sum(f.stat().st_size for f in ins),
                              "outputs": "tmpfs" if args.big_out == "tmpfs" else "/dev/null, /dev/zero (gathered, not kept: the input alone fills this box's RAM)",
                              "output_bytes": sum(f.stat().st_size for f in outs) if args.big_out == "tmpfs" else None, "discarded_input_path": disc, "seconds": round(dt, 2),
                              "pairs_per_s": round(big / dt), **ph, "device_GiB_high_water": hw, "stdout": so, "generate_s": round(gen_s, 1)}), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()

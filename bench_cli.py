"""Files in, files out: the drop-in binary (fastq-dupaway_b200/host/fastq-dupaway) against the reference binary
(oracle/_ref/fastq-dupaway, the unmodified sources compiled by oracle/Makefile) on the SAME input file, both started
the way a user starts them.  This is the end-to-end number of the whole product - host ingest (plain / single-member gzip /
multi-member gzip / BGZF), pinned staging, H2D, kernels, D2H, output gather and file writes - next to bench.py's kernel-level lines.

    python bench_cli.py [--reads 10000000] [--ref-reads 2000000] [--formats plain,gz1,gzmm,bgzf] [--mode fast|tight]

The input is the synthetic stream of bench.py (device generator, 150 bp, 30 % duplicates) written to tmpfs.  The
reference is timed on a prefix (--ref-reads; it is single-threaded, ~0.5 M reads/s) and on plain input only.  One JSON
line per run; outputs are compared between the two binaries on the common prefix run (byte identity, --fast).
Wall clock around the process (time.perf_counter), page cache warm, outputs on tmpfs.
"""
from __future__ import annotations

import argparse
import gzip
import importlib
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
EXE = ROOT / "fastq-dupaway_b200" / "host" / "fastq-dupaway"
READ_LEN, SEED, DUP_PERMILLE, N_PERMILLE = 150, 1, 300, 0
REC = 22 + 2 * READ_LEN


def synth_file(path: Path, n_reads: int, mate: int = 1, workers: int = 4):
    """device generator -> host -> file; `workers` threads take every k-th stretch of 2 M reads and pwrite it where it
    belongs (the generator is counter-based, tmpfs page allocation is the slow part and runs in parallel this way).
    FQD_BENCH_CLI_CPU_SYNTH=1: the numpy twin of the generator, for trying this script where there is no GPU."""
    if os.environ.get("FQD_BENCH_CLI_CPU_SYNTH"):
        gen = importlib.import_module("bench_synth")
        step = 4_000_000
        with open(path, "wb") as f:
            for first in range(0, n_reads, step):
                f.write(gen.synth_fastq_cpu(first, min(step, n_reads - first), READ_LEN, mate, SEED, DUP_PERMILLE, N_PERMILLE))
        return
    import ctypes as C
    pkg = importlib.import_module("fastq-dupaway_b200")
    lib = pkg.load_library()
    step = 2_000_000
    workers = max(1, min(workers, (n_reads + step - 1) // step))
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)

    def work(k):
        buf = pkg.DeviceBuffer(step * REC)
        host = C.create_string_buffer(step * REC)
        view = memoryview(host)
        for first in range(k * step, n_reads, step * workers):
            cnt = min(step, n_reads - first)
            assert lib.fqd_synth_fastq(0, buf.ptr, first, cnt, READ_LEN, mate, SEED, DUP_PERMILLE, N_PERMILLE, 0) == 0
            assert lib.fqd_memcpy_d2h(0, host, C.c_void_p(buf.ptr), cnt * REC) == 0
            done = 0
            while done < cnt * REC:
                done += os.pwrite(fd, view[done: cnt * REC], first * REC + done)
        buf.free()
    try:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(work, range(workers)))
    finally:
        os.close(fd)


def gz_members(src: Path, dst: Path, piece: int, bgzf: bool):
    """multi-member gzip (16 MiB members) or BGZF (64 KiB members with the BC extra subfield), compressed in threads"""
    data = src.read_bytes()

    def one(pos):
        part = data[pos:pos + piece]
        if not bgzf:
            return gzip.compress(part, compresslevel=1)
        c = zlib.compressobj(1, zlib.DEFLATED, -15)
        body = c.compress(part) + c.flush()
        bsize = 12 + 6 + len(body) + 8
        return (b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
                + body + struct.pack("<II", zlib.crc32(part), len(part)))
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4)) as ex, open(dst, "wb") as f:
        for blob in ex.map(one, range(0, len(data), piece)):
            f.write(blob)


def gz_single_member(src: Path, dst: Path, piece: int = 8 << 20):
    """ONE gzip member, as `gzip` / pigz write it, built from pieces deflated in threads (each ends in a sync flush,
    i.e. byte aligned and not final - pigz's construction, without its dictionary priming)"""
    data = src.read_bytes()
    starts = list(range(0, len(data), piece))

    def one(pos):
        c = zlib.compressobj(1, zlib.DEFLATED, -15)
        body = c.compress(data[pos:pos + piece])
        return body + (c.flush(zlib.Z_FINISH) if pos == starts[-1] else c.flush(zlib.Z_SYNC_FLUSH))
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 4)) as ex, open(dst, "wb") as f:
        f.write(b"\x1f\x8b\x08\x00\0\0\0\0\x00\xff")
        for blob in ex.map(one, starts):
            f.write(blob)
        f.write(struct.pack("<II", zlib.crc32(data), len(data) & 0xFFFFFFFF))


def timed(cmd, cwd, trace=False):
    env = dict(os.environ, FQD_TRACE="1") if trace else None
    t0 = time.perf_counter()
    res = subprocess.run([str(c) for c in cmd], cwd=cwd, capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    assert res.returncode == 0, res.stderr
    if not trace:
        return dt, res.stdout.strip()
    # [host-trace] <ms> ms  <label>: CUDA start-up (context, pinned staging, engine) vs the streaming part
    marks = {}
    extra = {}
    for line in res.stderr.splitlines():
        if line.startswith("[host-trace] device memory high-water mark"):
            extra["device_GiB_high_water"] = float(line.split("mark", 1)[1].split("GiB", 1)[0])
        elif line.startswith("[host-trace]"):
            ms, label = line[len("[host-trace]"):].split("ms", 1)
            label = label.strip()
            marks[label] = float(ms)
            for short in ("engine created", "outputs closed", "input on the device", "sorted / joined / scanned"):      # "... (raw input not kept on the device)"
                if label.startswith(short):
                    marks[short] = float(ms)
            if "raw input not kept" in label:
                extra["raw_input"] = "discarded after the parse (host gathers the output)"
    phases = {}
    if "engine created" in marks and "outputs closed" in marks:
        phases = {"startup_s": round(marks["engine created"] / 1e3, 3),
                  "stream_s": round((marks["outputs closed"] - marks["engine created"]) / 1e3, 3)}
        if "input on the device" in marks and "sorted / joined / scanned" in marks:
            phases["ingest_s"] = round((marks["input on the device"] - marks["engine created"]) / 1e3, 3)
            phases["sort_scan_s"] = round((marks["sorted / joined / scanned"] - marks["input on the device"]) / 1e3, 3)
            phases["output_s"] = round((marks["outputs closed"] - marks["sorted / joined / scanned"]) / 1e3, 3)
    phases.update(extra)
    return dt, res.stdout.strip(), phases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads (pairs with --paired)")
    ap.add_argument("--ref-reads", type=int, default=2_000_000)
    ap.add_argument("--formats", default="plain,gz1,gzmm,bgzf")
    ap.add_argument("--mode", default="fast", choices=["fast", "tight", "loose", "tail-hamming"])
    ap.add_argument("--paired", action="store_true", help="2 x 150 bp: -i/-u inputs, -o/-p outputs")
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--no-reference", action="store_true", help="skip the common-prefix runs of both binaries")
    args = ap.parse_args()
    oracle = importlib.import_module("oracle")
    tmp = Path(tempfile.mkdtemp(prefix="fqd_cli_", dir="/dev/shm" if Path("/dev/shm").is_dir() else None))
    mode_args = ["--fast"] if args.mode == "fast" else ["--compare-seq", args.mode]
    mates = (1, 2) if args.paired else (1,)
    unit = "pairs" if args.paired else "reads"

    def io_args(ins, outs):
        a = ["-i", ins[0], "-o", outs[0]]
        if args.paired:
            a += ["-u", ins[1], "-p", outs[1]]
        return a

    def make(fmt, src, dst):
        if fmt == "gzmm":
            gz_members(src, dst, 16 << 20, False)
        elif fmt == "bgzf":
            gz_members(src, dst, 0xff00, True)
        elif fmt == "gz1":
            gz_single_member(src, dst)
    try:
        full = [tmp / f"in_{m}.fq" for m in mates]
        for m, f in zip(mates, full):
            synth_file(f, args.reads, m)
        inputs = {"plain": full}
        for fmt in args.formats.split(","):
            if fmt != "plain":
                inputs[fmt] = [tmp / f"in_{fmt}_{m}.fq.gz" for m in mates]
                for src, dst in zip(full, inputs[fmt]):
                    make(fmt, src, dst)
        outs = [tmp / f"out_{m}.fq" for m in mates]
        for fmt in args.formats.split(","):
            best = None
            phases = {}
            for _ in range(args.repeats):
                dt, so, ph = timed([EXE, *io_args(inputs[fmt], outs), "-v", *mode_args], tmp, trace=True)
                if best is None or dt < best:
                    best, phases = dt, ph
            if phases.get("stream_s"):
                phases[f"stream_{unit}_per_s"] = round(args.reads / phases["stream_s"])
            if os.environ.get("FQD_WHOLE_INPUT"):
                phases["FQD_WHOLE_INPUT"] = os.environ["FQD_WHOLE_INPUT"]
            print(json.dumps({"impl": "ours", **phases, "binary": "fastq-dupaway_b200/host/fastq-dupaway", "mode": args.mode,
                              "paired": args.paired, "input": fmt, unit: args.reads,
                              "input_bytes": sum(f.stat().st_size for f in inputs[fmt]), "seconds": round(best, 3),
                              f"{unit}_per_s": round(args.reads / best),
                              "raw_GBps": round(args.reads * REC * len(mates) / best / 1e9, 3),
                              "io_threads": os.environ.get("FQD_IO_THREADS", "auto"), "host_cores": os.cpu_count(), "stdout": so}), flush=True)
        # common prefix: both binaries, outputs compared.  Sequence-based modes: the reference's choice inside a group of
        # equal records depends on its unstable sort, so the comparison is with the stable-sort build of the
        # same sources and the timing with the plain build.
        if args.no_reference:
            return
        n = min(args.ref_reads, args.reads)
        pre = [tmp / f"pre_{m}.fq" for m in mates]
        for src, dst in zip(full, pre):
            with open(src, "rb") as f, open(dst, "wb") as g:
                g.write(f.read(n * REC))
        o_ours = [tmp / f"o_ours_{m}.fq" for m in mates]
        o_ref = [tmp / f"o_ref_{m}.fq" for m in mates]
        dt_o, so_o = timed([EXE, *io_args(pre, o_ours), "-v", *mode_args], tmp)
        if oracle.ref_available():
            ref_args = mode_args + ([] if args.mode == "fast" else ["-m", "10240"])
            dt_r, so_r = timed([oracle.REF_BIN, *io_args(pre, o_ref), "-v", *ref_args], tmp)
            cmp_bin = "oracle/_ref/fastq-dupaway"
            if args.mode != "fast" and oracle.ref_available(stable=True):
                timed([oracle.REF_STABLE_BIN, *io_args(pre, o_ref), "-v", *ref_args], tmp)
                cmp_bin = "oracle/_ref/fastq-dupaway-stable"
            same = all(a.read_bytes() == b.read_bytes() for a, b in zip(o_ours, o_ref))
            print(json.dumps({"impl": "reference", "binary": "oracle/_ref/fastq-dupaway", "mode": args.mode, "paired": args.paired,
                              "input": "plain", unit: n, "seconds": round(dt_r, 3), f"{unit}_per_s": round(n / dt_r), "cores": 1,
                              "stdout": so_r, "ours_same_input_seconds": round(dt_o, 3), "outputs_byte_identical": same,
                              "outputs_compared_with": cmp_bin, "verbose_lines_identical": so_o == so_r}), flush=True)
            if "gz1" in inputs:
                # the reference on a single-member .gz of the same prefix (its gzip filter runs on the same one thread)
                pre_gz = [tmp / f"pre_{m}.fq.gz" for m in mates]
                for src, dst in zip(pre, pre_gz):
                    gz_single_member(src, dst)
                dt_g, _ = timed([oracle.REF_BIN, *io_args(pre_gz, o_ref), "-v", *ref_args], tmp)
                dt_og, _ = timed([EXE, *io_args(pre_gz, o_ours), "-v", *mode_args], tmp)
                print(json.dumps({"impl": "reference", "binary": "oracle/_ref/fastq-dupaway", "mode": args.mode, "paired": args.paired,
                                  "input": "gz1", unit: n, "seconds": round(dt_g, 3), f"{unit}_per_s": round(n / dt_g), "cores": 1,
                                  "ours_same_input_seconds": round(dt_og, 3)}), flush=True)
        else:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fastq-dupaway not built"}), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()

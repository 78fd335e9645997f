#!/usr/bin/env python
"""bench.py - throughput of the deduplication hot paths on B200 (contract in the build prompt).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation (oracle/_ref)

Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[1] - synthetic 100 M x 150 bp single-end
FASTQ, 30 % exact duplicates (32.2 GB, far larger than the 126 MB L2), `--fast`.  A "step" is one whole job: the hash set
is emptied, then every read goes through parse+pack -> insert -> duplicate count -> survivor index list, with the raw
FASTQ already resident in HBM (`value`), or pushed from pinned host memory through fqd_push_* with the per-chunk
results copied back (`e2e`).  In the same run the first 3 M reads go through the unmodified reference binary
(`cpu_baseline`) and our output on them must be byte-identical (`parity`).
`modes`: the metric's own shape, 2 x 150 bp paired-end, through `--fast`, every `--compare-seq` mode and `--fast
--unordered` (bench_modes.py) - each with its own value / roofline / e2e / cpu_baseline / parity.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

READ_LEN = 150
REC_BYTES = 22 + 2 * READ_LEN            # 322
SEED = 1
DUP_PERMILLE = 300
N_PERMILLE = 1
# algorithmic bytes per read (DESIGN.md "Algorithmic bytes"): K1 parse+pack reads the record once and writes
# the 64-byte key row, the 8-byte hash, the 4-byte record offset and the 1-byte flag
K1_BYTES_PER_READ = REC_BYTES + 64 + 8 + 4 + 1
METRIC = "dedup reads/sec"
UNIT = "reads/s"


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        sm = sorted(int(float(s[1])) for s in self.samples if len(s) > 2 and s[1].replace(".", "").isdigit())
        mx = [int(float(s[2])) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, nm in enumerate(names):
                if len(s) > 5 + k and s[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's own CPU implementation of the path (unmodified sources compiled into oracle/_ref, else the
    oracle port) on the host cores of this box.  The reference is single-threaded: cores = 1."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "oracle"))
    oracle = importlib.import_module("oracle")
    kind = "reference" if oracle.ref_available() else "port"
    gen = importlib.import_module("bench_synth")
    budget_s = float(os.environ.get("FQD_REF_BUDGET_S", 150))
    # calibrate on a small sample, then size the per-step sample so warmup+steps fit the budget
    tmp = Path(tempfile.mkdtemp(prefix="fqd_ref_", dir="/dev/shm" if Path("/dev/shm").is_dir() else None))
    try:
        def run(n_reads):
            buf = gen.synth_fastq_cpu(0, n_reads, READ_LEN, 1, SEED, DUP_PERMILLE, N_PERMILLE)
            t0 = time.perf_counter()
            if kind == "reference":
                rc, _, _, so, se = oracle.run_ref(tmp / "w", "fast", oracle.FASTQ, buf)
                assert rc == 0, se
            else:
                oracle.fast_se(buf, oracle.FASTQ)
            return time.perf_counter() - t0
        t_cal = run(100_000)
        rate = 100_000 / t_cal
        n_step = int(max(50_000, min(3_000_000, rate * budget_s / max(1, args.steps + args.warmup))))
        buf = gen.synth_fastq_cpu(0, n_step, READ_LEN, 1, SEED, DUP_PERMILLE, N_PERMILLE)
        inp = tmp / "in.fq"
        inp.write_bytes(buf)
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            if kind == "reference":
                res = subprocess.run([str(oracle.REF_BIN), "-i", str(inp), "-o", str(tmp / "out.fq"), "--fast"],
                                     cwd=tmp, capture_output=True)
                assert res.returncode == 0, res.stderr.decode()
            else:
                oracle.fast_se(buf, oracle.FASTQ)
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
        total = sum(times)
        value = n_step * len(times) / total
        sample = f"first {n_step} reads of the synthetic stream per step, plain FASTQ on tmpfs, --fast, wall clock around the process"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": workload_config(args, n_step),
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
                "sample_reads_per_step": n_step, "host_cores": os.cpu_count(),
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "modes": reference_modes(oracle, tmp) if kind == "reference" and not os.environ.get("FQD_BENCH_SKIP_MODES") else None}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_modes(oracle, tmp):
    """The reference binary on the metric's own shape (2 x 150 bp paired-end), one run per mode on a 1 M-pair prefix of the
    CPU twin of the generator (exact duplicates; the loose / tail-hamming variants of the stream exist on the device
    only - bench.py's own arm times the reference on those, next to its parity check)."""
    gen = importlib.import_module("bench_synth")
    n = int(os.environ.get("FQD_REF_MODE_PAIRS", 1_000_000))
    (tmp / "m1.fq").write_bytes(gen.synth_fastq_cpu(0, n, READ_LEN, 1, 2, DUP_PERMILLE, N_PERMILLE))
    (tmp / "m2.fq").write_bytes(gen.synth_fastq_cpu(0, n, READ_LEN, 2, 2, DUP_PERMILLE, N_PERMILLE))
    out = {}
    for mode, flags in (("fast_pe", ["--fast"]), ("tight", ["--compare-seq", "tight"]), ("loose", ["--compare-seq", "loose"]),
                        ("tail-hamming", ["--compare-seq", "tail-hamming", "--distance", "2"]), ("unordered", ["--fast", "--unordered"])):
        t0 = time.perf_counter()
        r = subprocess.run([str(oracle.REF_BIN), "-i", "m1.fq", "-u", "m2.fq", "-o", "mo1.fq", "-p", "mo2.fq", "-m", "10240", *flags], cwd=tmp, capture_output=True)
        dt = time.perf_counter() - t0
        out[mode] = {"value": n / dt if r.returncode == 0 else None, "unit": "pairs/s", "cores": 1, "kind": "reference",
                     "sample": f"first {n} pairs (exact duplicates), plain FASTQ on tmpfs, {' '.join(flags)} -m 10240, one run, wall clock"}
    return out


def workload_config(args, n_reads=None):
    """Identical in both arms.  The reference arm cannot take the whole workload inside a bounded run (0.5 M reads/s: 200 s
    per 100 M-read step), so it times the first reads of the SAME stream and reports a rate; how many is in its
    `cpu_baseline.sample` / `sample_reads_per_step`."""
    n_workload = int(os.environ.get("FQD_BENCH_READS", 100_000_000))
    return {"workload": "BASELINE configs[1]: synthetic 100Mx150bp single-end FASTQ, 30% exact duplicates, --fast",
            "reads_per_step": n_workload, "read_len": READ_LEN, "record_bytes": REC_BYTES, "dup_fraction": DUP_PERMILLE / 1000,
            "input": "larger than L2 (no flush needed)", "seed": SEED,
            "reference_arm": "a prefix of the same stream per step (bounded run; reads/s is a rate - the smaller set favours the reference's unordered_set)"}


# ---------------------------------------------------------------------------------------------------------
def cpu_baseline_and_parity(fqd, lib, dev, n_sample=3_000_000):
    """First n_sample reads of the workload (downloaded from the device generator): the unmodified reference binary on one
    core, wall clock (`cpu_baseline`), and this engine's output on the same bytes, which must be identical (`parity`)."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "oracle"))
    oracle = importlib.import_module("oracle")
    kind = "reference" if oracle.ref_available() else "port"
    tmp = Path(tempfile.mkdtemp(prefix="fqd_cpu_", dir="/dev/shm" if Path("/dev/shm").is_dir() else None))
    try:
        nbytes = n_sample * REC_BYTES
        dbuf = fqd.DeviceBuffer(nbytes + 65536, dev)
        for first in range(0, n_sample, 1_000_000):
            cnt = min(1_000_000, n_sample - first)
            assert lib.fqd_synth_fastq(dev, dbuf.ptr + first * REC_BYTES, first, cnt, READ_LEN, 1, SEED, DUP_PERMILLE, N_PERMILLE, 0) == 0
        buf = dbuf.download(nbytes)
        eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, False, False, 2, READ_LEN, n_sample + 1024, nbytes + 65536, n_sample + 1024, dev)
        eng.keep_survivors(True)
        res = eng.push_device(dbuf.ptr, nbytes)
        assert res.n_records == n_sample
        idx, _, _ = eng.survivors()
        st = eng.stats()
        eng.close()
        dbuf.free()
        ours = np.frombuffer(buf, dtype=np.uint8).reshape(n_sample, REC_BYTES)[idx.astype(np.int64)].tobytes()
        inp = tmp / "in.fq"
        inp.write_bytes(buf)
        t0 = time.perf_counter()
        if kind == "reference":
            r = subprocess.run([str(oracle.REF_BIN), "-i", str(inp), "-o", str(tmp / "out.fq"), "--fast", "-v"], cwd=tmp, capture_output=True)
            assert r.returncode == 0, r.stderr.decode()
            dt = time.perf_counter() - t0
            exp, exp_line = (tmp / "out.fq").read_bytes(), r.stdout.decode()
        else:
            eidx, est = oracle.fast_se(buf, oracle.FASTQ)
            dt = time.perf_counter() - t0
            exp = np.frombuffer(buf, dtype=np.uint8).reshape(n_sample, REC_BYTES)[eidx.astype(np.int64)].tobytes()
            exp_line = f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"
        line = f"{st.total} reads processed, out of which {st.dups} duplicates were removed.\n"
        parity = {"checked": True, "records": n_sample, "against": "oracle/_ref/fastq-dupaway (unmodified reference sources)" if kind == "reference" else "oracle port",
                  "bytes_identical": ours == exp, "summary_line_identical": line == exp_line, "output_bytes": len(ours)}
        parity["ok"] = parity["bytes_identical"] and parity["summary_line_identical"]
        assert parity["ok"], parity
        cpu = {"value": n_sample / dt, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"first {n_sample} reads of the same synthetic stream ({nbytes / 1e9:.2f} GB plain FASTQ on tmpfs), --fast, one run, wall clock"}
        return cpu, parity
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def our_arm(args):
    rank, local_rank, world = dist_env()
    fqd = importlib.import_module("fastq-dupaway_b200")
    lib = fqd.load_library()            # raises when the CUDA library is missing: no fallback
    dist = None
    force_sharded = bool(os.environ.get("FQD_BENCH_FORCE_SHARDED")) and "RANK" in os.environ      # profiling aid: N = 1 through the sharded path
    if world > 1 or force_sharded:
        import torch
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    n_total = int(os.environ.get("FQD_BENCH_READS", 100_000_000))
    if world > 1 or force_sharded:
        sys.path.insert(0, str(ROOT))
        mg = importlib.import_module("bench_multi")
        return mg.run(args, fqd, dist, rank, local_rank, world, n_total)

    chunk_reads = 6_000_000
    n_chunks = (n_total + chunk_reads - 1) // chunk_reads
    raw = fqd.DeviceBuffer(n_total * REC_BYTES + 65536, dev)
    for c in range(n_chunks):
        first = c * chunk_reads
        cnt = min(chunk_reads, n_total - first)
        rc = lib.fqd_synth_fastq(dev, raw.ptr + first * REC_BYTES, first, cnt, READ_LEN, 1, SEED, DUP_PERMILLE, N_PERMILLE, 0)
        assert rc == 0
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, False, False, 2, READ_LEN, n_total + 1024, chunk_reads * REC_BYTES + 65536,
                     chunk_reads + 1024, dev)
    eng.keep_survivors(True)         # device time = first kernel start -> survivor list ready

    def step_device():
        eng.reset()
        for c in range(n_chunks):
            first = c * chunk_reads
            cnt = min(chunk_reads, n_total - first)
            eng.push_device_async(raw.ptr + first * REC_BYTES, cnt * REC_BYTES)

    # ---- device-resident throughput (`value`)
    for _ in range(max(3, args.warmup)):
        step_device()
    eng.sync()
    st = eng.stats()
    assert st.err == 0 and st.total == n_total, (st.err, st.total)
    dups = st.dups
    sampler = ClockSampler(dev)
    sampler.start()
    eng.profile_enable(True)
    _, l0 = eng.device_time_ms()
    eng.timer_start()
    for _ in range(args.steps):
        step_device()
    ms_total = eng.timer_stop()
    eng.sync()
    _, l1 = eng.device_time_ms()
    prof = eng.profile()
    eng.profile_enable(False)
    sampler.stop_flag.set()
    sampler.join()
    st = eng.stats()
    assert st.err == 0 and st.total == n_total and st.dups == dups
    _, n_surv, _ = eng.survivors(fetch=False)
    assert n_surv == n_total - dups, (n_surv, n_total, dups)
    ms_per_step = ms_total / args.steps
    value = n_total / (ms_per_step / 1000.0)

    peak, peak_kind = measured_peak_gbs()
    k1_ms = prof.parse_ms / max(1, prof.parse_launches)
    reads_per_launch = (prof.parse_bytes / max(1, prof.parse_launches)) / REC_BYTES
    achieved = reads_per_launch * K1_BYTES_PER_READ / (k1_ms / 1000.0) / 1e9 if k1_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_parse_pack<4>", "achieved": achieved, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "alg_bytes_per_read": K1_BYTES_PER_READ, "reads_per_launch": reads_per_launch, "avg_launch_ms": k1_ms,
                "kernel_share_of_step": prof.parse_ms / ms_total if ms_total else None,
                "insert_share_of_step": prof.insert_ms / ms_total if ms_total else None,
                "whole_path_input_GBps": n_total * REC_BYTES / (ms_per_step / 1000.0) / 1e9,
                "whole_path": {"alg_bytes_per_read": 454, "alg_bytes_source": "SURVEY.md 8d (fast SE)", "this_build_bytes_per_read": 488 + 6,
                               "achieved": n_total * 454 / (ms_per_step / 1000.0) / 1e9, "frac": n_total * 454 / (ms_per_step / 1000.0) / 1e9 / peak,
                               "timed": "parse+pack, insert, duplicate count, survivor index list of every chunk of a job"}}
    tr = ROOT / "profiles" / "traffic.json"
    if tr.exists():
        try:
            # DRAM bytes per read of K1 from the committed ncu capture, scaled to this run's reads per launch
            roofline["traffic"] = json.loads(tr.read_text())["k_parse_pack_bytes_per_read"] * reads_per_launch
            roofline["traffic_source"] = "profiles/traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per read x reads per launch)"
        except Exception:
            pass

    # ---- end to end through the C ABI with HOST buffers (`e2e`)
    e2e = None
    try:
        if os.environ.get("FQD_BENCH_SKIP_E2E"):
            raise RuntimeError("skipped (FQD_BENCH_SKIP_E2E)")
        e2e = e2e_run(args, fqd, lib, eng, raw, n_total, chunk_reads, dups)
    except Exception as ex:   # report, never hide
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)}
    eng.close()
    raw.free()

    cpu, parity = (None, {"checked": False, "why": "FQD_BENCH_SKIP_CPU"}) if os.environ.get("FQD_BENCH_SKIP_CPU") else cpu_baseline_and_parity(fqd, lib, dev)
    modes = None
    if not os.environ.get("FQD_BENCH_SKIP_MODES"):
        bm = importlib.import_module("bench_modes")
        want = [m for m in os.environ.get("FQD_BENCH_MODES", "").split(",") if m] or None
        modes = bm.run_modes(fqd, lib, args, dev, (peak, peak_kind), want)
    files = None
    if not os.environ.get("FQD_BENCH_SKIP_FILES"):
        try:
            files = files_in_files_out()
        except Exception as ex:   # report, never hide
            files = {"error": repr(ex)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(args, n_total), "clocks": sampler.summary(),
            "e2e": e2e, "gpu_launches": int((l1 - l0)), "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "duplicates_removed": int(dups), "survivors_listed": int(n_surv), "input_GBps": n_total * REC_BYTES / (ms_per_step / 1000.0) / 1e9,
            "modes": modes, "files": files}
    print(json.dumps(line), flush=True)


def files_in_files_out():
    """What a user of the drop-in binary sees: fastq-dupaway_b200/host/fastq-dupaway started as a process on FILES (tmpfs),
    wall clock around it - CUDA start-up, host ingest, H2D, kernels, output gather and file writes included.  Bounded
    samples of the same synthetic stream: FQD_BENCH_FILE_READS single-end reads through --fast (default 20 M) and
    FQD_BENCH_FILE_PAIRS pairs through --compare-seq tight (default 10 M).  The reference's rate on files is `cpu_baseline`."""
    bc = importlib.import_module("bench_cli")
    shm = Path("/dev/shm")
    n_se = int(os.environ.get("FQD_BENCH_FILE_READS", 20_000_000))
    n_pe = int(os.environ.get("FQD_BENCH_FILE_PAIRS", 10_000_000))
    need = max(n_se, 2 * n_pe) * REC_BYTES * 2
    if not shm.is_dir() or shutil.disk_usage(shm).free < need * 1.5:
        return {"error": "not enough room on /dev/shm for the sample files"}
    if not bc.EXE.exists():
        subprocess.run(["make", "-s", "-C", str(bc.EXE.parent)], check=True)
    tmp = Path(tempfile.mkdtemp(prefix="fqd_files_", dir=shm))
    out = {}
    try:
        a, b = tmp / "in_1.fq", tmp / "in_2.fq"
        bc.synth_file(a, n_se, 1)
        best = None
        for _ in range(2):
            dt, so, ph = bc.timed([bc.EXE, "-i", a, "-o", tmp / "out_1.fq", "--fast", "-v"], tmp, trace=True)
            if best is None or dt < best[0]:
                best = (dt, so, ph)
        out["fast_se"] = {"value": n_se / best[0], "unit": "reads/s", "reads": n_se, "seconds": best[0], **best[2],
                          "input_bytes": a.stat().st_size, "output_bytes": (tmp / "out_1.fq").stat().st_size, "stdout": best[1]}
        (tmp / "out_1.fq").unlink()
        if n_pe:
            if n_pe != n_se:
                bc.synth_file(a, n_pe, 1)
            bc.synth_file(b, n_pe, 2)
            dt, so, ph = bc.timed([bc.EXE, "-i", a, "-u", b, "-o", tmp / "out_1.fq", "-p", tmp / "out_2.fq", "--compare-seq", "tight", "-v"], tmp, trace=True)
            out["tight_pe"] = {"value": n_pe / dt, "unit": "pairs/s", "pairs": n_pe, "seconds": dt, **ph, "input_bytes": a.stat().st_size + b.stat().st_size,
                               "output_bytes": (tmp / "out_1.fq").stat().st_size + (tmp / "out_2.fq").stat().st_size, "stdout": so}
        out["how"] = "plain FASTQ on tmpfs in, plain FASTQ on tmpfs out, one process per job, wall clock (time.perf_counter) around it; fast_se: best of 2"
        out["host_cores"] = os.cpu_count()
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def e2e_run(args, fqd, lib, eng, raw, n_total, chunk_reads, dups_expected):
    """Same job, inputs in pinned HOST memory, every chunk pushed with fqd_push (H2D inside) and its record
    offsets + duplicate flags read back (D2H inside)."""
    import ctypes as C
    n_e2e = int(os.environ.get("FQD_BENCH_E2E_READS", n_total))
    chunk = min(chunk_reads, 3_000_000)
    nbytes = n_e2e * REC_BYTES
    host = C.c_void_p()
    rc = lib.fqd_host_alloc(C.byref(host), nbytes)
    if rc:
        raise MemoryError(f"pinned host allocation of {nbytes} bytes failed")
    try:
        # fill the pinned buffer from the device-resident synthetic data (same bytes as the `value` run)
        step = 1 << 30
        for o in range(0, nbytes, step):
            n = min(step, nbytes - o)
            assert lib.fqd_memcpy_d2h(eng.cfg.device, C.c_void_p(host.value + o), C.c_void_p(raw.ptr + o), n) == 0
        n_chunks = (n_e2e + chunk - 1) // chunk
        d2h = 0

        def one_step():
            nonlocal d2h
            eng.reset()
            surv = 0
            d2h = 0
            def span(c):
                first = c * chunk
                return first, min(chunk, n_e2e - first)
            # fqd_push split in two: the H2D copy of chunk c+1 runs while chunk c is processed and its results come back
            f0, c0 = span(0)
            assert lib.fqd_push_prefetch(eng.h, C.c_void_p(host.value + f0 * REC_BYTES), c0 * REC_BYTES, None, 0) == 0
            for c in range(n_chunks):
                first, cnt = span(c)
                if c + 1 < n_chunks:
                    f1, c1 = span(c + 1)
                    assert lib.fqd_push_prefetch(eng.h, C.c_void_p(host.value + f1 * REC_BYTES), c1 * REC_BYTES, None, 0) == 0
                res = fqd.ChunkResult()
                rc2 = lib.fqd_push_staged(eng.h, C.byref(res))
                assert rc2 == 0 and res.n_records == cnt
                surv += res.n_survivors
                d2h += (cnt + 1) * 4 + cnt + 64
            return surv
        for _ in range(1):
            one_step()
        steps = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(steps):
            surv = one_step()
        dt = (time.perf_counter() - t0) / steps
        if n_e2e == n_total:
            assert n_e2e - surv == dups_expected
        return {"value": n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h,
                "reads_per_step": n_e2e, "ms_per_step": dt * 1000.0, "steps": steps,
                "path": "fqd_push_prefetch / fqd_push_staged (pinned host FASTQ -> H2D of chunk c+1 overlapping kernels + D2H of record "
                        "offsets + duplicate flags of chunk c), wall clock"}
    finally:
        lib.fqd_host_free(host)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        our_arm(args)


if __name__ == "__main__":
    main()

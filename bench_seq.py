#!/usr/bin/env python
"""bench_seq.py - throughput of the whole-input modes on one B200 (BASELINE.json configs[2] and configs[3]).

  python bench_seq.py --mode tight|loose|tail-hamming|unordered [--pairs N] [--steps K] [--cpu]

Workload: synthetic 2x150 bp paired-end FASTQ (644 B per pair), 30 % duplicates, the SURVEY section 8d variants
(loose: 10 % of the duplicates truncated; tail-hamming: 10 % of the duplicates with <= 2 tail substitutions;
unordered: the same pairs with R2 read in a different order).  A step = one whole job: the engine adopts the
input that is resident in HBM (fqd_adopt_device: parse + pack in place), then fqd_finish runs sort + scan (or tag
sort + join + pair set).  The synthetic input is generated on the device outside the timed region; device time is
the CUDA-event interval around adopt + finish.  Prints one JSON line per mode; profiles/ keeps the committed results.
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
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

READ_LEN = 150
REC = 22 + 2 * READ_LEN
SEED = 2
DUP_PERMILLE = 300
N_PERMILLE = 1
VARIANT = {"tight": 0, "loose": 1, "tail-hamming": 2, "unordered": 0}
# algorithmic bytes per pair: both records read once + 2 x 64 B key rows written (K1),
# (8 B word + 4 B index) read twice and written once per 8-bit radix pass over the first key word (8 passes), the
# two key rows of a pair and of its predecessor read by the scan, 4 B of survivor list
ALG_BYTES_PER_PAIR = 2 * REC + 2 * 64 + 8 * 36 + 2 * 128 + 4


def run_mode(fqd, lib, mode, n_pairs, steps, dev=0):
    unordered = mode == "unordered"
    emode = "fast" if unordered else mode
    chunk_pairs = 1_500_000
    n_chunks = (n_pairs + chunk_pairs - 1) // chunk_pairs
    # the whole input of both mates is generated into HBM once (outside the timed region); the engine adopts the
    # buffers (fqd_adopt_device: parsed in place, no copy)
    raw = [fqd.DeviceBuffer(n_pairs * REC + 65536, dev) for _ in range(2)]
    off2 = 0
    for c in range(n_chunks):
        first = c * chunk_pairs
        cnt = min(chunk_pairs, n_pairs - first)
        assert lib.fqd_synth_fastq(dev, raw[0].ptr + first * REC, first, cnt, READ_LEN, 1, SEED, DUP_PERMILLE, N_PERMILLE, VARIANT[mode]) == 0
        # unordered: R2 arrives chunk-reversed (its tag sort has to undo that; pairs are matched by tag)
        c2 = (n_chunks - 1 - c) if unordered else c
        first2 = c2 * chunk_pairs
        cnt2 = min(chunk_pairs, n_pairs - first2)
        assert lib.fqd_synth_fastq(dev, raw[1].ptr + off2, first2, cnt2, READ_LEN, 2, SEED, DUP_PERMILLE, N_PERMILLE, VARIANT[mode]) == 0
        off2 += cnt2 * REC
    eng = fqd.Engine(emode, fqd.FORMAT_FASTQ, True, unordered, 2, READ_LEN, n_pairs + 1024, 1 << 30, 0, dev, 16)
    times = []
    stats = None
    for it in range(steps + 1):
        if it:
            eng.reset()
        eng.timer_start()
        eng.adopt_device(0, raw[0].ptr, n_pairs * REC)
        eng.adopt_device(1, raw[1].ptr, n_pairs * REC)
        eng.finish()
        t_ms = eng.timer_stop()
        st = eng.stats()
        assert st.err == 0, (st.err, st.err_record)
        if it:
            times.append(t_ms)
        stats = st
    em = eng.emission()
    n_out = int(em.n_out)
    _, launches = eng.device_time_ms()
    eng.close()
    for s in raw:
        s.free()
    ms = sum(times) / len(times)
    peak = 6538.6
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text())["hbm_gbs"])
    line = {"metric": "dedup read pairs/sec", "mode": mode, "value": n_pairs / (ms / 1e3), "unit": "pairs/s",
            "reads_per_s": 2 * n_pairs / (ms / 1e3), "n_gpus": 1, "steps": steps, "ms_per_step": ms,
            "config": {"workload": f"synthetic {n_pairs} x 2x150bp paired-end FASTQ, 30% duplicates, "
                                   + ("--fast --unordered (R2 chunk-reversed)" if unordered else f"--compare-seq {mode}"),
                       "pairs": n_pairs, "record_bytes": REC, "seed": SEED},
            "pairs_total": int(stats.total), "duplicates_removed": int(stats.dups), "unmatched": int(stats.unmatched),
            "pairs_out": n_out, "input_GBps": n_pairs * 2 * REC / (ms / 1e3) / 1e9,
            "alg_GBps": n_pairs * ALG_BYTES_PER_PAIR / (ms / 1e3) / 1e9, "alg_bytes_per_pair": ALG_BYTES_PER_PAIR,
            "frac_of_measured_peak": n_pairs * ALG_BYTES_PER_PAIR / (ms / 1e3) / 1e9 / peak}
    return line


def run_mode_multi(fqd, lib, mode, n_pairs_per_rank, steps, emit=True):
    """N > 1 (under torchrun): weak scaling, every rank contributes n_pairs_per_rank pairs of ONE global stream
    (rank r holds pairs [r * n, (r + 1) * n)); key-range sharding through fastq-dupaway_b200/sharded_seq.py."""
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    sh = importlib.import_module("fastq-dupaway_b200.sharded_seq")
    dev = local
    chunk_pairs = 1_500_000
    n = n_pairs_per_rank
    n_chunks = (n + chunk_pairs - 1) // chunk_pairs
    # this rank's slice of both mates, generated into HBM once (outside the timed region) and adopted in place
    raw = [fqd.DeviceBuffer(n * REC + 65536, dev) for _ in range(2)]
    for c in range(n_chunks):
        first = c * chunk_pairs
        cnt = min(chunk_pairs, n - first)
        for m in range(2):
            assert lib.fqd_synth_fastq(dev, raw[m].ptr + first * REC, rank * n + first, cnt, READ_LEN, m + 1, SEED, DUP_PERMILLE, N_PERMILLE,
                                       VARIANT[mode]) == 0
    ops = sh.GpuRangeOps(fqd, mode, fqd.FORMAT_FASTQ, True, 2, READ_LEN, n + 1024, int(n * 1.3) + (1 << 20), dev, seg_bytes=1 << 30)
    peer = None
    if not os.environ.get("FQD_NO_PEER"):
        px = importlib.import_module("fastq-dupaway_b200.peer")
        peer = [px.PeerExchange(fqd, dist, rank, world, dev, int(n * REC * 1.3) + (64 << 20)) for _ in range(2)]     # one receive buffer per mate
    times = []
    res = None
    for it in range(steps + 1):
        if it:
            ops.reset()
        if it == 1:
            sh.TRACE.clear()          # FQD_TRACE: the warm-up job (pool growth, NCCL set-up) is not representative
        torch.cuda.synchronize(dev)
        ops.origin.timer_start()
        for m in range(2):
            ops.origin.adopt_device(m, raw[m].ptr, n * REC)
        t_append = ops.origin.timer_stop()              # parse + pack of this rank's slice (same as the N = 1 arm)
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = sh.dedup_ranges(ops, dist, rank, world, n_samples=8192, peer=peer)
        torch.cuda.synchronize(dev)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1) + t_append], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if it:
            times.append(float(ms.item()))
    tot = torch.tensor(list(res), dtype=torch.int64, device=f"cuda:{dev}")
    own = [torch.zeros_like(tot) for _ in range(world)]
    dist.all_gather(own, tot)
    line = {}
    if rank == 0:
        ms = sum(times) / len(times)
        total = world * n
        owned = [int(t[0].item()) for t in own]
        line = {"metric": "dedup read pairs/sec", "mode": mode, "value": total / (ms / 1e3), "unit": "pairs/s",
                "reads_per_s": 2 * total / (ms / 1e3), "n_gpus": world, "steps": steps, "ms_per_step": ms, "scaling": "weak",
                "config": {"workload": f"synthetic {total} x 2x150bp paired-end FASTQ, 30% duplicates, --compare-seq {mode}",
                           "pairs_per_gpu": n, "parallelism": f"key-range x{world}", "record_bytes": REC, "seed": SEED,
                           "timed": "adopt (parse + pack of the slice in place) + sample + splitters + plan + gather + all-to-all of raw records + parse + sort + scan + "
                                    "boundary chain + emission lists (slices already appended and resident)"},
                "pairs_total": total, "duplicates_removed": int(sum(int(t[2].item()) for t in own)),
                "pairs_out": int(sum(int(t[1].item()) for t in own)), "owned_per_rank": owned,
                "imbalance": max(owned) / (sum(owned) / world), "input_GBps": total * 2 * REC / (ms / 1e3) / 1e9,
                "alltoall_bytes_per_gpu": n * 2 * REC,
                "exchange": "mapped peer memory (CUDA IPC + copy engines)" if peer is not None else "NCCL all_to_all_single"}
        if emit:
            print(json.dumps(line), flush=True)
        if sh.TRACE:
            print("[fqd trace] ms over all steps:", json.dumps({k: round(v, 1) for k, v in sh.TRACE.items()}), file=sys.stderr)
            sh.TRACE.clear()
    ops.close()
    if peer is not None:
        for px_ in peer:
            px_.close()
    for s_ in raw:
        s_.free()
    return line


def run_unordered_multi(fqd, lib, n_pairs_per_rank, steps, seed=SEED, emit=True):
    """N > 1 (under torchrun): --fast --unordered over tag ranges (fastq-dupaway_b200/sharded_unordered.py), weak scaling.
    ONE global job of world * n pairs, dealt to the ranks chunk by chunk round-robin: file 1 in input order, file 2 with
    its chunks in reverse order, so every rank holds records of every tag range in both files (a true all-to-all) and
    no rank holds the mates of its own file-1 records."""
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    sh = importlib.import_module("fastq-dupaway_b200.sharded_unordered")
    dev = local
    n = n_pairs_per_rank
    chunk_pairs = 1_250_000
    n_chunks = (n + chunk_pairs - 1) // chunk_pairs
    assert n % chunk_pairs == 0, "pairs per GPU must be a multiple of 1.25 M"
    total_chunks = n_chunks * world
    raw = [fqd.DeviceBuffer(n * REC + 65536, dev) for _ in range(2)]
    for c in range(n_chunks):
        g1 = c * world + rank                       # global chunk of file 1
        g2 = total_chunks - 1 - g1                  # file 2: the chunks arrive in reverse order
        for m, g in ((0, g1), (1, g2)):
            assert lib.fqd_synth_fastq(dev, raw[m].ptr + c * chunk_pairs * REC, g * chunk_pairs, chunk_pairs, READ_LEN, m + 1, seed,
                                       DUP_PERMILLE, N_PERMILLE, 0) == 0
    ops = sh.GpuTagRangeOps(fqd, fqd.FORMAT_FASTQ, READ_LEN, n + 1024, int(n * 1.3) + (1 << 20), dev, seg_bytes=1 << 30, max_tag_len=16)
    times = []
    res = None
    for it in range(steps + 1):
        if it:
            ops.reset()
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for m in range(2):
            ops.adopt(m, raw[m].ptr, n * REC)       # parse + pack + tags of this rank's slices, in place
        res = sh.dedup_tag_ranges(ops, dist, rank, world, n_samples=8192)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if it:
            times.append(float(ms.item()))
    own = torch.tensor([res.local_pairs, res.local_kept], dtype=torch.int64, device=f"cuda:{dev}")
    allown = [torch.zeros_like(own) for _ in range(world)]
    dist.all_gather(allown, own)
    line = {}
    if rank == 0:
        ms = sum(times) / len(times)
        total = world * n
        owned = [int(t[0].item()) for t in allown]
        line = {"metric": "dedup read pairs/sec", "mode": "unordered", "value": total / (ms / 1e3), "unit": "pairs/s",
                "reads_per_s": 2 * total / (ms / 1e3), "n_gpus": world, "steps": steps, "ms_per_step": ms, "scaling": "weak",
                "config": {"workload": f"synthetic {total} x 2x150bp paired-end FASTQ, 30% duplicates, --fast --unordered "
                                       "(file 2 chunk-reversed; chunks of 1.25 M pairs dealt round-robin to the ranks)",
                           "pairs_per_gpu": n, "parallelism": f"tag-range x{world} (join) + hash-range x{world} (pair set)", "record_bytes": REC,
                           "seed": seed,
                           "timed": "adopt (parse + pack + tags of the slices in place) + sample + splitters + plan + gather + all-to-all of the raw "
                                    "records of both files + parse + tag sorts + join + stop state + all-to-all of the pair keys + pair set + "
                                    "flags back + emission lists (slices resident in HBM)"},
                "err": res.err, "pairs_total": res.total, "duplicates_removed": res.dups, "unmatched": res.unmatched,
                "pairs_out": res.total - res.dups, "pairs_per_rank": owned, "imbalance": max(owned) / max(1.0, sum(owned) / world),
                "input_GBps": total * 2 * REC / (ms / 1e3) / 1e9,
                "alltoall_bytes_per_gpu": n * 2 * REC + n * ops.row_bytes, "exchange": "NCCL all_to_all_single"}
        if emit:
            print(json.dumps(line), flush=True)
    ops.close()
    for s_ in raw:
        s_.free()
    return line


def cpu_reference(mode, n_pairs):
    """The unmodified reference (oracle/_ref) on the same synthetic stream, one core, tmpfs."""
    sys.path.insert(0, str(ROOT / "oracle"))
    oracle = importlib.import_module("oracle")
    if not oracle.ref_available():
        return None
    gen = importlib.import_module("bench_synth")
    if VARIANT[mode] != 0:
        return None          # the CPU twin of the generator only knows variant 0
    tmp = Path(tempfile.mkdtemp(prefix="fqd_cpu_", dir="/dev/shm" if Path("/dev/shm").is_dir() else None))
    try:
        (tmp / "r1.fq").write_bytes(gen.synth_fastq_cpu(0, n_pairs, READ_LEN, 1, SEED, DUP_PERMILLE, N_PERMILLE))
        (tmp / "r2.fq").write_bytes(gen.synth_fastq_cpu(0, n_pairs, READ_LEN, 2, SEED, DUP_PERMILLE, N_PERMILLE))
        cmd = [str(oracle.REF_BIN), "-i", "r1.fq", "-u", "r2.fq", "-o", "o1.fq", "-p", "o2.fq", "-m", "10240"]
        if mode == "unordered":
            cmd += ["--fast", "--unordered"]
        else:
            cmd += ["--compare-seq", mode]
        t0 = time.perf_counter()
        res = subprocess.run(cmd, cwd=tmp, capture_output=True)
        dt = time.perf_counter() - t0
        assert res.returncode == 0, res.stderr.decode()
        return {"value": n_pairs / dt, "unit": "pairs/s", "cores": 1, "kind": "reference",
                "sample": f"{n_pairs} pairs of the same synthetic stream, plain FASTQ on tmpfs, -m 10240, wall clock"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="all")
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("FQD_BENCH_PAIRS", 20_000_000)))
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--cpu", action="store_true", help="also time the reference binary on a 1 M-pair sample")
    args = ap.parse_args()
    fqd = importlib.import_module("fastq-dupaway_b200")
    lib = fqd.load_library()
    if int(os.environ.get("WORLD_SIZE", 1)) > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        modes = ["tight", "loose", "tail-hamming"] if args.mode == "all" else [args.mode]
        for m in modes:
            run_mode_multi(fqd, lib, m, args.pairs, args.steps)
        dist.barrier()
        dist.destroy_process_group()
        return
    modes = ["tight", "loose", "tail-hamming", "unordered"] if args.mode == "all" else [args.mode]
    for m in modes:
        line = run_mode(fqd, lib, m, args.pairs, args.steps)
        if args.cpu:
            line["cpu_baseline"] = cpu_reference(m, 1_000_000)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

"""bench.py's N > 1 path: deduplication of ONE global read stream sharded over the N GPUs of one box, one process per GPU.

Headline (`value`): `--fast` on single-end reads, hash-range sharding (fastq-dupaway_b200/sharded2.py, csrc/shard2.cuh).
Weak scaling: every rank contributes FQD_BENCH_READS reads (default 100 M x 150 bp = 32.2 GB in its HBM); the global
stream is N times that and duplicates reference ANY earlier read of the global stream, so the exchange is real.  Timed on
the device (CUDA events over both streams of every rank), max over ranks.  Outside the timed region rank 0 pushes the
same global stream through ONE single-GPU engine and the duplicate counts must agree (`verify`).
`modes`: the metric's own shape (2 x 150 bp paired-end) - `--fast` through the same sharded path, and every
`--compare-seq` mode through key-range sharding (fastq-dupaway_b200/sharded_seq.py, bench_seq.py)."""
from __future__ import annotations

import importlib
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent


def _sharded_fast(args, fqd, lib, sharded2, dist, gloo, rank, dev, world, n_per_rank, paired, seed, steps, warmup, b):
    """One sharded --fast workload; returns (ms_per_step max over ranks, duplicates of the global stream, profile, launches, engine bits)."""
    import torch
    REC = b.REC_BYTES
    chunk_reads = int(os.environ.get("FQD_BENCH_CHUNK_READS", 10_000_000 if not paired else 5_000_000))
    n_chunks = (n_per_rank + chunk_reads - 1) // chunk_reads
    mates = 2 if paired else 1
    raw = [fqd.DeviceBuffer(n_per_rank * REC + 65536, dev) for _ in range(mates)]
    # chunk c of rank r = global records [(c*world + r) * chunk_reads, ...): chunks are dealt round-robin in global input order
    sizes = []
    for c in range(n_chunks):
        cnt = min(chunk_reads, n_per_rank - c * chunk_reads)
        first_global = (c * world + rank) * chunk_reads
        for m in range(mates):
            assert lib.fqd_synth_fastq(dev, raw[m].ptr + c * chunk_reads * REC, first_global, cnt, b.READ_LEN, m + 1, seed, b.DUP_PERMILLE, b.N_PERMILLE, 0) == 0
        sizes.append(cnt)
    region = sharded2.region_rows_for(chunk_reads, world)
    region = (region + 15) // 16 * 16
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, paired, False, 2, b.READ_LEN, n_chunks * world * region + 1024,
                     chunk_reads * REC + 65536, chunk_reads + 1024, dev)
    ops = sharded2.GpuShard2Ops(fqd, eng, world, rank, region)
    sharded2.connect(ops, dist, rank, world)
    chunks = [tuple(x for m in range(mates) for x in (raw[m].ptr + c * chunk_reads * REC, sizes[c] * REC)) for c in range(n_chunks)]

    barrier = gloo          # a callable: the shared-memory host barrier

    def step():
        barrier()
        ops.reset()
        barrier()
        return sharded2.run_job(ops, barrier, chunks)

    for _ in range(warmup):
        total, dups = step()
    assert total == n_per_rank
    barrier()
    eng.profile_enable(True)
    _, l0 = eng.device_time_ms()
    ops.timer_start()
    for _ in range(steps):
        total2, dups2 = step()
    ms_local = ops.timer_stop()
    barrier()
    assert (total2, dups2) == (total, dups)
    ms = torch.tensor([ms_local], dtype=torch.float64, device=f"cuda:{dev}")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot = torch.tensor([dups, total], dtype=torch.int64, device=f"cuda:{dev}")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    _, l1 = eng.device_time_ms()
    prof = eng.profile()
    eng.profile_enable(False)
    res = {"ms_per_step": float(ms.item()) / steps, "dups": int(tot[0].item()), "total": int(tot[1].item()), "prof": prof, "launches": int(l1 - l0),
           "chunk_reads": chunk_reads, "region_rows": region, "row_bytes": 64 * mates + 8, "ms_total": float(ms.item())}
    barrier()
    eng.close()
    for r in raw:
        r.free()
    barrier()
    return res


def _single_gpu_duplicates(fqd, lib, dev, world, n_per_rank, chunk_reads, paired, seed, b):
    """The same global stream through ONE engine on this GPU (generated chunk by chunk into a scratch buffer)."""
    REC = b.REC_BYTES
    mates = 2 if paired else 1
    n_total = world * n_per_rank
    scratch = [fqd.DeviceBuffer(chunk_reads * REC + 65536, dev) for _ in range(mates)]
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, paired, False, 2, b.READ_LEN, n_total + 1024, chunk_reads * REC + 65536, chunk_reads + 1024, dev)
    n_chunks_rank = (n_per_rank + chunk_reads - 1) // chunk_reads
    for c in range(n_chunks_rank):
        cnt = min(chunk_reads, n_per_rank - c * chunk_reads)
        for r in range(world):                               # global input order: chunk c of rank 0, 1, ...
            first_global = (c * world + r) * chunk_reads
            for m in range(mates):
                assert lib.fqd_synth_fastq(dev, scratch[m].ptr, first_global, cnt, b.READ_LEN, m + 1, seed, b.DUP_PERMILLE, b.N_PERMILLE, 0) == 0
            res = eng.push_device(scratch[0].ptr, cnt * REC, scratch[1].ptr if paired else None, cnt * REC if paired else 0)
            assert res.n_records == cnt
    st = eng.stats()
    assert st.err == 0 and st.total == n_total, (st.err, st.total)
    eng.close()
    for s_ in scratch:
        s_.free()
    return int(st.dups)


def run(args, fqd, dist, rank, local_rank, world, n_per_rank):
    import torch
    b = importlib.import_module("bench")
    sharded2 = importlib.import_module("fastq-dupaway_b200.sharded2")
    lib = fqd.load_library()
    dev = local_rank
    torch.cuda.set_device(dev)
    gloo = sharded2.HostBarrier(fqd, dist, rank, world)   # host barriers: shared memory + two atomics, nothing is launched on the GPUs for them
    REC = b.REC_BYTES
    sampler = b.ClockSampler(dev)
    if rank == 0:
        sampler.start()
    r = _sharded_fast(args, fqd, lib, sharded2, dist, gloo, rank, dev, world, n_per_rank, False, b.SEED, args.steps, max(3, args.warmup), b)
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()
    verify = None
    if not os.environ.get("FQD_BENCH_SKIP_VERIFY"):
        if rank == 0:
            try:
                single = _single_gpu_duplicates(fqd, lib, dev, world, n_per_rank, r["chunk_reads"], False, b.SEED, b)
                verify = {"single_gpu_duplicates": single, "sharded_duplicates": r["dups"], "equal": single == r["dups"],
                          "how": "rank 0 pushed the same global stream (chunks in global input order) through one single-GPU engine, outside the timed region"}
            except Exception as ex:
                verify = {"equal": None, "error": repr(ex)}
        gloo()
        if rank == 0 and verify.get("equal") is False:
            raise AssertionError(f"sharded duplicate count differs from the single-GPU run: {verify}")

    modes = None
    if not os.environ.get("FQD_BENCH_SKIP_MODES"):
        modes = {}
        n_pairs = int(os.environ.get("FQD_BENCH_PAIRS_PER_GPU", 25_000_000))        # configs[2]: 200 M pairs at N = 8
        msteps = max(1, min(args.steps, int(os.environ.get("FQD_BENCH_MODE_STEPS", 2))))
        peak, peak_kind = b.measured_peak_gbs()
        t0 = time.perf_counter()
        pe = _sharded_fast(args, fqd, lib, sharded2, dist, gloo, rank, dev, world, n_pairs, True, 2, msteps, 1, b)
        pv = None
        if rank == 0 and not os.environ.get("FQD_BENCH_SKIP_VERIFY"):
            try:
                single = _single_gpu_duplicates(fqd, lib, dev, world, n_pairs, pe["chunk_reads"], True, 2, b)
                pv = {"single_gpu_duplicates": single, "sharded_duplicates": pe["dups"], "equal": single == pe["dups"]}
            except Exception as ex:
                pv = {"equal": None, "error": repr(ex)}
        gloo()
        if rank == 0:
            assert pv is None or pv.get("equal") is not False, pv
            ach = pe["total"] * 860 / (pe["ms_per_step"] / 1e3) / 1e9
            modes["fast_pe"] = {"value": pe["total"] / (pe["ms_per_step"] / 1e3), "unit": "pairs/s", "reads_per_s": 2 * pe["total"] / (pe["ms_per_step"] / 1e3),
                                "ms_per_step": pe["ms_per_step"], "steps": msteps, "pairs_per_gpu": n_pairs, "pairs_total": pe["total"],
                                "duplicates_removed": pe["dups"], "verify": pv, "gpu_launches": pe["launches"],
                                "roofline": {"bound": "hbm", "kernel": "whole path", "achieved": ach, "peak": peak * world, "peak_kind": peak_kind, "unit": "GB/s",
                                             "frac": ach / (peak * world), "alg_bytes_per_pair": 860, "alg_bytes_source": "SURVEY.md 8d"},
                                "parallelism": f"hash-range x{world}", "leg_wall_s": round(time.perf_counter() - t0, 1)}
        bs = importlib.import_module("bench_seq")
        for mode in ("tight", "loose", "tail-hamming"):
            t0 = time.perf_counter()
            try:
                line = bs.run_mode_multi(fqd, lib, mode, int(os.environ.get("FQD_BENCH_SEQ_PAIRS_PER_GPU", 20_000_000)), msteps, emit=False)
            except Exception as ex:
                line = {"value": None, "unit": "pairs/s", "error": repr(ex)}
            if rank == 0:
                if line.get("value"):
                    ach = line["pairs_total"] * 1068 / (line["ms_per_step"] / 1e3) / 1e9
                    line["roofline"] = {"bound": "hbm", "kernel": "whole path", "achieved": ach, "peak": peak * world, "peak_kind": peak_kind, "unit": "GB/s",
                                        "frac": ach / (peak * world), "alg_bytes_per_pair": 1068, "alg_bytes_source": "SURVEY.md 8d"}
                line["leg_wall_s"] = round(time.perf_counter() - t0, 1)
                modes[mode] = line
        t0 = time.perf_counter()
        try:
            # the pairs of the fast_pe leg (same generator, seed and count) with file 2 in another order: same duplicates
            line = bs.run_unordered_multi(fqd, lib, n_pairs, msteps, seed=2, emit=False)
        except Exception as ex:
            line = {"value": None, "unit": "pairs/s", "error": repr(ex)}
        if rank == 0:
            if line.get("value"):
                ach = line["pairs_total"] * 1156 / (line["ms_per_step"] / 1e3) / 1e9
                line["roofline"] = {"bound": "hbm", "kernel": "whole path", "achieved": ach, "peak": peak * world, "peak_kind": peak_kind, "unit": "GB/s",
                                    "frac": ach / (peak * world), "alg_bytes_per_pair": 1156, "alg_bytes_source": "bench_modes.py (SURVEY.md 8d + tag rows)"}
                uv = {"pairs_total": line["pairs_total"], "expected_pairs": world * n_pairs, "unmatched": line["unmatched"],
                      "duplicates_removed": line["duplicates_removed"], "fast_pe_duplicates_same_pairs": pe["dups"],
                      "how": "same pairs as the fast_pe leg (whose duplicate count is checked against one single-GPU engine); every pair must be "
                             "matched, none skipped, and the same number removed"}
                uv["equal"] = (line["err"] == 0 and uv["pairs_total"] == uv["expected_pairs"] and uv["unmatched"] == 0
                               and uv["duplicates_removed"] == pe["dups"])
                line["verify"] = uv
                assert uv["equal"], uv
            line["leg_wall_s"] = round(time.perf_counter() - t0, 1)
            modes["unordered"] = line

    if rank == 0:
        prof = r["prof"]
        ms_per_step = r["ms_per_step"]
        n_total = r["total"]
        value = n_total / (ms_per_step / 1000.0)
        peak, peak_kind = b.measured_peak_gbs()
        k1_ms = prof.parse_ms / max(1, prof.parse_launches)
        rpl = (prof.parse_bytes / max(1, prof.parse_launches)) / REC
        achieved = rpl * b.K1_BYTES_PER_READ / (k1_ms / 1000.0) / 1e9 if k1_ms > 0 else 0.0
        line = {"metric": b.METRIC, "value": value, "unit": b.UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": dict(b.workload_config(args), parallelism=f"hash-range x{world}",
                                                     reads_per_gpu=n_per_rank, chunk_reads=r["chunk_reads"]),
                "clocks": sampler.summary(),
                "e2e": {"value": None, "unit": b.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "end-to-end (host buffers) is measured at N=1; the N>1 arm measures the sharded device path"},
                "gpu_launches": r["launches"],
                "roofline": {"bound": "hbm", "kernel": "k_parse_pack<4>", "achieved": achieved, "peak": peak, "peak_kind": peak_kind,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": None, "alg_bytes_per_read": b.K1_BYTES_PER_READ,
                             "avg_launch_ms": k1_ms, "kernel_share_of_step": prof.parse_ms / r["ms_total"],
                             "insert_share_of_step": prof.insert_ms / r["ms_total"],
                             "scatter_share_of_step": prof.scatter_ms / r["ms_total"],
                             "note": "rank 0's K1 launches, timed while the other stream of the same GPU inserts the previous chunk"},
                "exchange": {"row_bytes": r["row_bytes"], "bytes_over_nvlink_per_gpu_per_step": int(n_per_rank * (r["row_bytes"] + 1) * (world - 1) / world),
                             "region_rows": r["region_rows"], "host_barriers_per_chunk": 1,
                             "rows": "written by the scatter kernel straight into the owners' key-store regions over mapped peer memory (CUDA IPC, NVLink stores "
                                     "from the SMs); flags written back the same way; ordering by interprocess CUDA events"},
                "verify": verify, "duplicates_removed": r["dups"], "input_GBps": n_total * REC / (ms_per_step / 1000.0) / 1e9, "modes": modes}
        print(json.dumps(line), flush=True)
    gloo()
    gloo.close()
    dist.destroy_process_group()

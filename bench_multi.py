"""bench.py's N > 1 path: --fast deduplication of ONE global read stream sharded by hash range over N GPUs.
Weak scaling: every rank contributes FQD_BENCH_READS reads (default 100 M x 150 bp = 32.2 GB in its HBM); the global
stream is N times that, and duplicates reference ANY earlier read of the global stream, so the all-to-all is real."""
from __future__ import annotations

import importlib
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent


def run(args, fqd, dist, rank, local_rank, world, n_per_rank):
    import torch
    b = importlib.import_module("bench")
    sharded = importlib.import_module("fastq-dupaway_b200.sharded")
    lib = fqd.load_library()
    dev = local_rank
    torch.cuda.set_device(dev)
    REC = b.REC_BYTES
    chunk_reads = int(os.environ.get("FQD_BENCH_CHUNK_READS", 10_000_000))     # < 4 GiB per chunk (u32 offsets); fewer, larger chunks amortise the per-chunk collectives
    n_chunks = (n_per_rank + chunk_reads - 1) // chunk_reads
    raw = fqd.DeviceBuffer(n_per_rank * REC + 65536, dev)
    # chunk c of rank r = global reads [(c*world + r) * chunk_reads, ...): chunks are fed in global input order
    sizes = []
    for c in range(n_chunks):
        cnt = min(chunk_reads, n_per_rank - c * chunk_reads)
        first_global = (c * world + rank) * chunk_reads
        rc = lib.fqd_synth_fastq(dev, raw.ptr + c * chunk_reads * REC, first_global, cnt, b.READ_LEN, 1, b.SEED, b.DUP_PERMILLE, b.N_PERMILLE, 0)
        assert rc == 0
        sizes.append(cnt)
    # every rank owns 1/world of the key space: ~n_per_rank rows arrive here (+ imbalance margin)
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, False, False, 2, b.READ_LEN, int(n_per_rank * 1.15) + (1 << 20),
                     chunk_reads * REC + 65536, chunk_reads + 1024, dev)
    ops = sharded.GpuShardOps(fqd, eng, world, dev, chunk_reads + 1024)
    peer = None
    packers, peers2 = None, None
    pipelined = not os.environ.get("FQD_NO_PEER") and not os.environ.get("FQD_NO_PIPELINE")
    if not os.environ.get("FQD_NO_PEER"):
        px = importlib.import_module("fastq-dupaway_b200.peer")
        cap = int((chunk_reads + 1024) * ops.row_bytes * 1.25) + (16 << 20)
        if pipelined:
            # two pack-only engines on their own streams work one chunk ahead of the exchange; two receive buffers
            peers2 = [px.PeerExchange(fqd, dist, rank, world, dev, cap) for _ in range(2)]
            pengs = [fqd.Engine("fast", fqd.FORMAT_FASTQ, False, False, 2, b.READ_LEN, 1 << 16, chunk_reads * REC + 65536,
                                chunk_reads + 1024, dev) for _ in range(2)]
            packers = [sharded.GpuShardOps(fqd, e, world, dev, chunk_reads + 1024, own_stream=True) for e in pengs]
            ops.pack(raw.ptr, 0)                      # sets the main engine up for fqd_shard_insert
        else:
            peer = px.PeerExchange(fqd, dist, rank, world, dev, cap)
    chunks = [(raw.ptr + c * chunk_reads * REC, sizes[c] * REC) for c in range(n_chunks)]

    def step():
        eng.reset()
        if pipelined:
            return sharded.exchange_pipelined(packers, ops, dist, world, chunks, peers2)
        dups = 0
        for c in range(n_chunks):
            d, _ = sharded.exchange_chunk(ops, dist, world, chunks[c][0], chunks[c][1], peer=peer)
            dups += d
        return dups

    for _ in range(max(3, args.warmup)):
        dups = step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = b.ClockSampler(dev)
    if rank == 0:
        sampler.start()
    eng.profile_enable(True)
    if pipelined:
        for e in pengs:
            e.profile_enable(True)
    _, l0 = eng.device_time_ms()
    if pipelined:
        l0 += sum(e.device_time_ms()[1] for e in pengs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(args.steps):
        d2 = step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    assert d2 == dups
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{dev}")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tot = torch.tensor([dups, n_per_rank], dtype=torch.int64, device=f"cuda:{dev}")
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    _, l1 = eng.device_time_ms()
    if pipelined:
        l1 += sum(e.device_time_ms()[1] for e in pengs)
    prof = eng.profile()
    if pipelined:                                     # K1 runs on the pack engines, K2 on the main one
        for e in pengs:
            pp = e.profile()
            prof.parse_ms += pp.parse_ms; prof.parse_launches += pp.parse_launches; prof.parse_bytes += pp.parse_bytes
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join()
        ms_per_step = float(ms.item()) / args.steps
        n_total = int(tot[1].item())
        value = n_total / (ms_per_step / 1000.0)
        peak, peak_kind = b.measured_peak_gbs()
        k1_ms = prof.parse_ms / max(1, prof.parse_launches)
        rpl = (prof.parse_bytes / max(1, prof.parse_launches)) / REC
        achieved = rpl * b.K1_BYTES_PER_READ / (k1_ms / 1000.0) / 1e9 if k1_ms > 0 else 0.0
        row_bytes = ops.row_bytes
        line = {"metric": b.METRIC, "value": value, "unit": b.UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": dict(b.workload_config(args, n_total), parallelism=f"hash-range x{world}",
                                                     reads_per_gpu=n_per_rank, chunk_reads=chunk_reads),
                "clocks": sampler.summary(),
                "e2e": {"value": None, "unit": b.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "end-to-end (host buffers) is measured at N=1; the N>1 arm measures the sharded device path"},
                "gpu_launches": int(l1 - l0),
                "roofline": {"bound": "hbm", "kernel": "k_parse_pack<4>", "achieved": achieved, "peak": peak, "peak_kind": peak_kind,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": None, "alg_bytes_per_read": b.K1_BYTES_PER_READ,
                             "avg_launch_ms": k1_ms, "kernel_share_of_step": prof.parse_ms / float(ms.item()),
                             "insert_share_of_step": prof.insert_ms / float(ms.item())},
                "exchange": {"row_bytes": row_bytes, "alltoall_bytes_per_gpu_per_step": n_per_rank * (row_bytes + 1),
                             "collectives_per_chunk": 3,
                             "rows": ("mapped peer memory (CUDA IPC + copy engines), overlapped with the split + pack of the next chunk" if pipelined
                                      else "mapped peer memory (CUDA IPC + copy engines)" if peer is not None else "NCCL all_to_all_single")},
                "duplicates_removed": int(tot[0].item()), "input_GBps": n_total * REC / (ms_per_step / 1000.0) / 1e9}
        print(json.dumps(line), flush=True)
    if rank == 0 and sharded.TRACE:
        print("[fqd trace] per-phase wall clock, ms over all steps:", json.dumps({k: round(v, 2) for k, v in sharded.TRACE.items()}), file=sys.stderr)
    eng.close()
    if peer is not None:
        peer.close()
    if pipelined:
        for e in pengs:
            e.close()
        for p_ in peers2:
            p_.close()
    raw.free()
    dist.barrier()
    dist.destroy_process_group()

"""GPU test of the round-2 sharded --fast path (csrc/shard2.cuh, fastq-dupaway_b200/sharded2.py): 2 and 3 ranks as separate
processes that share cuda:0 - rows written straight into the owners' key-store regions over mapped (CUDA IPC) memory,
ordering by interprocess events, flags written back into the sources' flag regions.  The union of the ranks' duplicate
flags must equal the oracle's decision on the GLOBAL stream (first occurrence in global input order survives), single-
and paired-end, over several chunks (both parities of the double-buffered regions are reused) and two jobs per handle."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, chunks, counts, result_dir, jobs):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fqd = importlib.import_module("fastq-dupaway_b200")
    sharded2 = importlib.import_module("fastq-dupaway_b200.sharded2")
    mine = chunks[rank]
    paired = isinstance(mine[0], tuple)
    maxb = (max(max(len(a), len(b)) for a, b in mine) if paired else max(len(c) for c in mine)) + 4096
    per_chunk = max(counts)
    region = sharded2.region_rows_for(per_chunk, world) + per_chunk // 4          # small chunks: generous
    n_chunks = len(mine)
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, paired, False, 2, 100, n_chunks * world * ((region + 15) // 16 * 16) + 1024, maxb, per_chunk + 1024, 0)
    ops = sharded2.GpuShard2Ops(fqd, eng, world, rank, region)
    sharded2.connect(ops, dist, rank, world)
    bufs = []
    for c in mine:
        if paired:
            b1, b2 = fqd.DeviceBuffer(maxb, 0), fqd.DeviceBuffer(maxb, 0)
            b1.upload(c[0]); b2.upload(c[1])
            bufs.append((b1.ptr, len(c[0]), b2.ptr, len(c[1]), b1, b2))
        else:
            b = fqd.DeviceBuffer(maxb, 0)
            b.upload(c)
            bufs.append((b.ptr, len(c), b))
    args = [t[:4] if paired else t[:2] for t in bufs]
    for job in range(jobs):
        if job:
            dist.barrier()
            ops.reset()
            dist.barrier()
        flags = []
        total, dups = sharded2.run_job(ops, dist.barrier, args, flags_out=flags, records=counts)
        f = np.concatenate([np.frombuffer(x, dtype=np.uint8) for x in flags])
        assert total == sum(counts) and dups == int(f.sum())
        np.save(Path(result_dir) / f"flags_{job}_{rank}.npy", f)
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def _check(tmp_path, world, n_chunks, per_chunk, exp, jobs):
    for job in range(jobs):
        got = np.zeros(len(exp), dtype=np.uint8)
        for r in range(world):
            f = np.load(tmp_path / f"flags_{job}_{r}.npy")
            for c in range(n_chunks):
                lo = (c * world + r) * per_chunk
                got[lo: lo + per_chunk] = f[c * per_chunk: (c + 1) * per_chunk]
        assert np.array_equal(got, exp), f"job {job}: {int((got != exp).sum())} flags differ"


@pytest.mark.parametrize("world", [2, 3])
def test_regions_single_end(tmp_path, oracle, world):
    n_chunks, per_chunk = 5, 6000
    seqs = synth.make_reads(world * n_chunks * per_chunk, seed=81 + world, read_len=100, var_len=True, n_frac=0.02, dup_frac=0.4)
    recs = [synth.to_fastq([s], ids=[b"@g.%d" % i]) for i, s in enumerate(seqs)]
    # chunk c of rank r = global records [(c*world + r) * per_chunk, ...)
    chunks = [[b"".join(recs[(c * world + r) * per_chunk: (c * world + r + 1) * per_chunk]) for c in range(n_chunks)] for r in range(world)]
    port = 29800 + (os.getpid() * 3 + world) % 2000
    mp.spawn(_worker, args=(world, port, chunks, [per_chunk] * n_chunks, str(tmp_path), 2), nprocs=world, join=True)
    keep_idx, est = oracle.fast_se(b"".join(recs), oracle.FASTQ)
    exp = np.ones(len(recs), dtype=np.uint8)
    exp[keep_idx.astype(np.int64)] = 0
    _check(tmp_path, world, n_chunks, per_chunk, exp, 2)


def test_regions_paired_end(tmp_path, oracle):
    """Paired-end: the rows carry both mates' keys (128 bytes), ownership follows the hash of the pair."""
    world, n_chunks, per_chunk = 2, 4, 3000
    s1, s2 = synth.make_pair(world * n_chunks * per_chunk, seed=83, read_len=80, var_len=True, n_frac=0.02, dup_frac=0.4)
    r1 = [synth.to_fastq([s], ids=[b"@g.%d 1" % i]) for i, s in enumerate(s1)]
    r2 = [synth.to_fastq([s], ids=[b"@g.%d 2" % i]) for i, s in enumerate(s2)]

    def cut(recs, c, r):
        return b"".join(recs[(c * world + r) * per_chunk: (c * world + r + 1) * per_chunk])
    chunks = [[(cut(r1, c, r), cut(r2, c, r)) for c in range(n_chunks)] for r in range(world)]
    port = 29900 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, chunks, [per_chunk] * n_chunks, str(tmp_path), 1), nprocs=world, join=True)
    keep_idx, est = oracle.fast_pe(b"".join(r1), b"".join(r2), oracle.FASTQ)
    exp = np.ones(len(r1), dtype=np.uint8)
    exp[keep_idx.astype(np.int64)] = 0
    _check(tmp_path, world, n_chunks, per_chunk, exp, 1)

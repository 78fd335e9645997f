"""GPU test of the sharded (multi-GPU) --fast kernels: two ranks, both on cuda:0, exchanging through gloo on the host
(NCCL refuses two ranks on one device).  The union of the ranks' duplicate flags must equal the oracle's decision on the
global stream."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, chunks, result_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fqd = importlib.import_module("fastq-dupaway_b200")
    sharded = importlib.import_module("fastq-dupaway_b200.sharded")
    torch.cuda.set_device(0)
    mine = chunks[rank]
    maxb = max(len(c) for c in mine) + 4096 if not isinstance(mine[0], tuple) else 0
    paired = isinstance(mine[0], tuple)
    if paired:
        maxb = max(max(len(a), len(b)) for a, b in mine) + 4096
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, paired, False, 2, 100, 200000, maxb, 50000, 0)
    ops = sharded.GpuShardOps(fqd, eng, world, 0, 50000)
    buf = fqd.DeviceBuffer(maxb, 0)
    buf2 = fqd.DeviceBuffer(maxb, 0) if paired else None
    flags = []
    for c in mine:
        if paired:
            buf.upload(c[0]); buf2.upload(c[1])
            d, n = sharded.exchange_chunk(ops, dist, world, buf.ptr, len(c[0]), via_cpu=True, raw2_ptr=buf2.ptr, nbytes2=len(c[1]))
        else:
            buf.upload(c)
            d, n = sharded.exchange_chunk(ops, dist, world, buf.ptr, len(c), via_cpu=True)
        res = np.frombuffer(C_string(eng, n), dtype=np.uint8).copy()
        assert int(res.sum()) == d
        flags.append(res)
    np.save(Path(result_dir) / f"flags_{rank}.npy", np.concatenate(flags))
    st = eng.stats()
    assert st.err == 0
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def C_string(eng, n):
    """duplicate flags of the last chunk (device -> host) through the C ABI helper"""
    import ctypes as C
    lib = eng.lib
    out = C.create_string_buffer(int(n))
    lib.fqd_shard_read_flags.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    assert lib.fqd_shard_read_flags(eng.h, out, int(n)) == 0
    return out.raw


def test_two_ranks_one_gpu(tmp_path, oracle):
    world, n_chunks, per_chunk = 2, 3, 4000
    seqs = synth.make_reads(world * n_chunks * per_chunk, seed=51, read_len=100, var_len=True, n_frac=0.02, dup_frac=0.4)
    recs = [synth.to_fastq([s], ids=[b"@g.%d" % i]) for i, s in enumerate(seqs)]
    # chunk c of rank r = global records [(c*world + r) * per_chunk, ...)
    chunks = [[b"".join(recs[(c * world + r) * per_chunk: (c * world + r + 1) * per_chunk]) for c in range(n_chunks)] for r in range(world)]
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, chunks, str(tmp_path)), nprocs=world, join=True)
    keep_idx, est = oracle.fast_se(b"".join(recs), oracle.FASTQ)
    exp = np.ones(len(recs), dtype=np.uint8)
    exp[keep_idx.astype(np.int64)] = 0
    got = np.zeros(len(recs), dtype=np.uint8)
    for r in range(world):
        f = np.load(tmp_path / f"flags_{r}.npy")
        for c in range(n_chunks):
            lo = (c * world + r) * per_chunk
            got[lo: lo + per_chunk] = f[c * per_chunk: (c + 1) * per_chunk]
    assert np.array_equal(got, exp)


def test_two_ranks_one_gpu_paired(tmp_path, oracle):
    """Paired-end: the exchanged rows carry both mates' keys, ownership follows the hash of the pair."""
    world, n_chunks, per_chunk = 2, 3, 3000
    s1, s2 = synth.make_pair(world * n_chunks * per_chunk, seed=52, read_len=80, var_len=True, n_frac=0.02, dup_frac=0.4)
    r1 = [synth.to_fastq([s], ids=[b"@g.%d 1" % i]) for i, s in enumerate(s1)]
    r2 = [synth.to_fastq([s], ids=[b"@g.%d 2" % i]) for i, s in enumerate(s2)]
    def cut(recs, c, r):
        return b"".join(recs[(c * world + r) * per_chunk: (c * world + r + 1) * per_chunk])
    chunks = [[(cut(r1, c, r), cut(r2, c, r)) for c in range(n_chunks)] for r in range(world)]
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, chunks, str(tmp_path)), nprocs=world, join=True)
    keep_idx, est = oracle.fast_pe(b"".join(r1), b"".join(r2), oracle.FASTQ)
    exp = np.ones(len(r1), dtype=np.uint8)
    exp[keep_idx.astype(np.int64)] = 0
    got = np.zeros(len(r1), dtype=np.uint8)
    for r in range(world):
        f = np.load(tmp_path / f"flags_{r}.npy")
        for c in range(n_chunks):
            lo = (c * world + r) * per_chunk
            got[lo: lo + per_chunk] = f[c * per_chunk: (c + 1) * per_chunk]
    assert np.array_equal(got, exp)


def _worker_pipelined(rank, world, port, chunks, result_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fqd = importlib.import_module("fastq-dupaway_b200")
    sharded = importlib.import_module("fastq-dupaway_b200.sharded")
    peer = importlib.import_module("fastq-dupaway_b200.peer")
    torch.cuda.set_device(0)
    mine = chunks[rank]
    maxb = max(len(c) for c in mine) + 4096
    main_eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, False, False, 2, 100, 200000, maxb, 50000, 0)
    main = sharded.GpuShardOps(fqd, main_eng, world, 0, 50000)
    pengs = [fqd.Engine("fast", fqd.FORMAT_FASTQ, False, False, 2, 100, 1024, maxb, 50000, 0) for _ in range(2)]
    packers = [sharded.GpuShardOps(fqd, e, world, 0, 50000, own_stream=True) for e in pengs]
    peers = [peer.PeerExchange(fqd, dist, rank, world, 0, 50000 * main.row_bytes) for _ in range(2)]
    bufs = []
    for c in mine:
        b = fqd.DeviceBuffer(len(c) + 64, 0)
        b.upload(c)
        bufs.append(b)
    main.pack(bufs[0].ptr, 0)                 # sets the main engine up for fqd_shard_insert
    flags = []
    dups = sharded.exchange_pipelined(packers, main, dist, world, [(b.ptr, len(c)) for b, c in zip(bufs, mine)], peers,
                                      via_cpu=True, flags_out=flags)
    got = np.concatenate([np.frombuffer(f, dtype=np.uint8) for f in flags])
    assert int(got.sum()) == dups
    np.save(Path(result_dir) / f"flags_{rank}.npy", got)
    dist.barrier()
    for p_ in peers:
        p_.close()
    for e in pengs + [main_eng]:
        e.close()
    dist.destroy_process_group()


def test_two_ranks_one_gpu_pipelined_peer_memory(tmp_path, oracle):
    """The pipelined exchange (pack-only engines one chunk ahead, key rows over mapped peer memory): per-record duplicate
    flags against the oracle's decision on the global stream."""
    world, n_chunks, per_chunk = 2, 4, 3000
    seqs = synth.make_reads(world * n_chunks * per_chunk, seed=53, read_len=100, var_len=True, n_frac=0.02, dup_frac=0.4)
    recs = [synth.to_fastq([s], ids=[b"@g.%d" % i]) for i, s in enumerate(seqs)]
    chunks = [[b"".join(recs[(c * world + r) * per_chunk: (c * world + r + 1) * per_chunk]) for c in range(n_chunks)] for r in range(world)]
    port = 29800 + (os.getpid() % 2000)
    mp.spawn(_worker_pipelined, args=(world, port, chunks, str(tmp_path)), nprocs=world, join=True)
    keep_idx, est = oracle.fast_se(b"".join(recs), oracle.FASTQ)
    exp = np.ones(len(recs), dtype=np.uint8)
    exp[keep_idx.astype(np.int64)] = 0
    got = np.zeros(len(recs), dtype=np.uint8)
    for r in range(world):
        f = np.load(tmp_path / f"flags_{r}.npy")
        for c in range(n_chunks):
            lo = (c * world + r) * per_chunk
            got[lo: lo + per_chunk] = f[c * per_chunk: (c + 1) * per_chunk]
    assert np.array_equal(got, exp)

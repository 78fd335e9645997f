"""CPU test (gloo ranks) of the host loop of the round-2 sharded --fast protocol (fastq-dupaway_b200/sharded2.py).

The device side is replaced by a numpy stand-in with the same contract, in which "peer memory" is a set of memory-mapped
files every rank can write: pack() scatters a chunk's keys straight into the owners' key-store REGIONS (one region per
(chunk, source)), insert() walks the regions of a chunk in source order, flags travel back through the sources' flag
regions.  A stand-in call executes when it is issued, so what is tested is exactly what the barriers of run_job must
guarantee on the GPU: an owner only looks at a chunk after EVERY source has issued its scatter, a source only reads flags
after every owner has issued them, and the double-buffered (chunk parity) regions are not reused before they were read."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class MapShard2Ops:
    def __init__(self, shared: Path, world, rank, region_rows, n_chunks):
        self.world, self.rank, self.R = world, rank, region_rows
        self.keys = [np.lib.format.open_memmap(shared / f"keys_{r}.npy", mode="r+") for r in range(world)]        # [n_chunks, world, R]
        self.counts = [np.lib.format.open_memmap(shared / f"counts_{r}.npy", mode="r+") for r in range(world)]    # [2, world]
        self.flags_in = [np.lib.format.open_memmap(shared / f"flags_{r}.npy", mode="r+") for r in range(world)]   # [2, world, R]
        self.seen = {}
        self.dest = [None, None]
        self.n = [0, 0]
        self.dup = None
        self.total = self.dups = 0

    @staticmethod
    def owner(keys, world):
        h = (keys * np.uint64(0x9E3779B97F4A7C15)) ^ (keys >> np.uint64(29))
        return ((h.astype(object) * world) >> 64).astype(np.int64)

    def pack(self, chunk, keys, _n=0):
        par = chunk & 1
        o = self.owner(keys, self.world)
        pos = np.zeros(len(keys), dtype=np.int64)
        for k in range(self.world):
            idx = np.flatnonzero(o == k)
            pos[idx] = np.arange(len(idx))
            assert len(idx) <= self.R
            self.keys[k][chunk, self.rank, : len(idx)] = keys[idx]
            self.counts[k][par, self.rank] = len(idx)
        self.dest[par] = (o, pos)
        self.n[par] = len(keys)
        self.total += len(keys)

    def insert(self, chunk):
        par = chunk & 1
        mine = self.keys[self.rank]
        for s in range(self.world):
            cnt = int(self.counts[self.rank][par, s])
            flags = np.zeros(cnt, dtype=np.uint8)
            for j in range(cnt):
                k = int(mine[chunk, s, j])
                if k in self.seen:
                    flags[j] = 1
                else:
                    self.seen[k] = (chunk, s, j)
            self.flags_in[s][par, self.rank, :cnt] = flags

    def apply(self, chunk):
        par = chunk & 1
        o, pos = self.dest[par]
        self.dup = self.flags_in[self.rank][par, o, pos].copy()
        self.dups += int(self.dup.sum())

    def read_flags(self, n):
        return self.dup[:n].tobytes()

    def finish(self):
        return self.total, self.dups


def _worker(rank, world, port, n_chunks, chunk, shared):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    sharded2 = importlib.import_module("fastq-dupaway_b200.sharded2")
    shared = Path(shared)
    R = 2 * chunk // world + 64
    if rank == 0:
        for r in range(world):
            np.lib.format.open_memmap(shared / f"keys_{r}.npy", mode="w+", dtype=np.uint64, shape=(n_chunks, world, R)).flush()
            np.lib.format.open_memmap(shared / f"counts_{r}.npy", mode="w+", dtype=np.int64, shape=(2, world)).flush()
            np.lib.format.open_memmap(shared / f"flags_{r}.npy", mode="w+", dtype=np.uint8, shape=(2, world, R)).flush()
    dist.barrier()
    rng = np.random.default_rng(321)
    stream = rng.integers(0, 900, size=world * n_chunks * chunk).astype(np.uint64)      # many duplicates
    ops = MapShard2Ops(shared, world, rank, R, n_chunks)
    chunks = [(stream[(c * world + rank) * chunk: (c * world + rank + 1) * chunk],) for c in range(n_chunks)]
    flags = []
    total, dups = sharded2.run_job(ops, dist.barrier, chunks, flags_out=flags, records=[chunk] * n_chunks)
    assert total == n_chunks * chunk
    np.save(shared / f"dups_{rank}.npy", np.concatenate([np.frombuffer(f, dtype=np.uint8) for f in flags]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_region_protocol(tmp_path, world):
    n_chunks, chunk = 5, 400
    port = 29500 + (os.getpid() * 7 + world) % 2000
    mp.spawn(_worker, args=(world, port, n_chunks, chunk, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(321)
    stream = rng.integers(0, 900, size=world * n_chunks * chunk).astype(np.uint64)
    seen, exp = set(), np.zeros(len(stream), dtype=np.uint8)          # first occurrence in GLOBAL input order survives
    for i, k in enumerate(stream):
        if int(k) in seen:
            exp[i] = 1
        else:
            seen.add(int(k))
    got = np.zeros(len(stream), dtype=np.uint8)
    for r in range(world):
        d = np.load(tmp_path / f"dups_{r}.npy")
        for c in range(n_chunks):
            lo = (c * world + r) * chunk
            got[lo: lo + chunk] = d[c * chunk: (c + 1) * chunk]
    assert np.array_equal(got, exp)


def test_region_rows_margin():
    import importlib
    sharded2 = importlib.import_module("fastq-dupaway_b200.sharded2")
    for chunk, world in ((10_000_000, 8), (10_000_000, 2), (1000, 3)):
        r = sharded2.region_rows_for(chunk, world)
        assert r > chunk / world and r < 1.2 * chunk / world + 8192


def _bar_worker(rank, world, port, shared):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    fqd = importlib.import_module("fastq-dupaway_b200")
    sharded2 = importlib.import_module("fastq-dupaway_b200.sharded2")
    bar = sharded2.HostBarrier(fqd, dist, rank, world)
    cnt = np.lib.format.open_memmap(Path(shared) / "cnt.npy", mode="r+")
    for it in range(300):
        cnt[rank] = it + 1
        bar()
        assert all(int(cnt[r]) >= it + 1 for r in range(world)), (it, list(cnt))      # nobody passes before everybody arrived
        bar()
    bar.close()
    dist.barrier()
    dist.destroy_process_group()


def test_shared_memory_host_barrier(tmp_path):
    world = 3
    np.lib.format.open_memmap(tmp_path / "cnt.npy", mode="w+", dtype=np.int64, shape=(world,)).flush()
    port = 29500 + (os.getpid() * 11 + 5) % 2000
    mp.spawn(_bar_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)

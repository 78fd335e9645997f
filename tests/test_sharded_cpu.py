"""CPU test (2 gloo ranks) of the multi-GPU exchange plumbing in fastq-dupaway_b200/sharded.py.  The device functions
pack / insert / apply are replaced by a small numpy stand-in with the same contract, so that what is tested here is
the host-side protocol: split sizes, ordering of the received rows, and the way duplicate flags travel back."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class NumpyShardOps:
    """Stand-in for GpuShardOps: rows = (8-byte key, 8-byte hash); owner = hash * world >> 64."""

    def __init__(self, world):
        self.world = world
        self.seen = set()
        self.order = None
        self.dups = None

    def pack(self, keys, _nbytes):
        h = (keys * np.uint64(0x9E3779B97F4A7C15)) ^ (keys >> np.uint64(31))
        owner = ((h.astype(object) * self.world) >> 64).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        self.order = order
        rows = np.stack([keys[order], h[order]], axis=1).astype(np.uint64)
        counts = [int((owner == k).sum()) for k in range(self.world)]
        return torch.from_numpy(rows.view(np.uint8).reshape(len(keys), 16).copy()), counts

    def insert(self, recv_rows):
        rows = recv_rows.numpy().view(np.uint64).reshape(-1, 2)
        flags = np.zeros(len(rows), dtype=np.uint8)
        for i, k in enumerate(rows[:, 0]):
            if int(k) in self.seen:
                flags[i] = 1
            else:
                self.seen.add(int(k))
        return torch.from_numpy(flags)

    def apply(self, back):
        f = back.numpy()
        d = np.zeros(len(f), dtype=np.uint8)
        d[self.order] = f
        self.dups = d
        return int(f.sum())


def _worker(rank, world, port, n_chunks, chunk, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    sharded = importlib.import_module("fastq-dupaway_b200.sharded")
    rng = np.random.default_rng(123)
    stream = rng.integers(0, 400, size=world * n_chunks * chunk).astype(np.uint64)      # many duplicates
    ops = NumpyShardOps(world)
    out = []
    for c in range(n_chunks):
        lo = (c * world + rank) * chunk
        keys = stream[lo: lo + chunk]
        d, n = sharded.exchange_chunk(ops, dist, world, keys, 0, via_cpu=True)
        assert n == chunk
        out.append(ops.dups.copy())
    np.save(Path(result_dir) / f"dups_{rank}.npy", np.concatenate(out))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_protocol_two_ranks(tmp_path):
    world, n_chunks, chunk = 2, 3, 500
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_chunks, chunk, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    stream = rng.integers(0, 400, size=world * n_chunks * chunk).astype(np.uint64)
    # reference: first occurrence in GLOBAL input order survives
    seen, exp = set(), np.zeros(len(stream), dtype=np.uint8)
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

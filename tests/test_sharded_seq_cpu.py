"""CPU test (3 gloo ranks) of the multi-GPU sequence-mode plumbing in fastq-dupaway_b200/sharded_seq.py.  The device
side (GpuRangeOps) is replaced by a numpy stand-in with the same contract whose "records" are 8-byte integers and
whose comparator is an adjacent rule that does NOT respect range ownership (two keys are duplicates when they fall
into the same decade), so the result is only right when splitters, split sizes, arrival order and the boundary
chain - including ranks that own nothing - all work."""
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class NumpyRangeOps:
    def __init__(self, keys):
        self.keys = np.asarray(keys, dtype=np.uint64)
        self.recv = []
        self.fixed = 0

    def sample(self, n):
        if len(self.keys) == 0:
            return np.full((n, 2), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64), 0
        idx = (np.arange(n) * len(self.keys)) // n
        return np.stack([self.keys[idx], np.zeros(n, dtype=np.uint64)], axis=1), len(self.keys)

    def plan(self, splitters, world):
        sp = [(int(a), int(b)) for a, b in splitters]
        owner = np.array([sum(1 for s in sp if s <= (int(k), 0)) for k in self.keys], dtype=np.int64)
        self.order = np.argsort(owner, kind="stable")
        counts = [int((owner == o).sum()) for o in range(world)]
        return counts, [[c * 8 for c in counts]]

    def gather(self, mate, total):
        return torch.from_numpy(self.keys[self.order].view(np.uint8).copy())

    def receive(self, mate, recv):
        self.recv.append(recv.numpy().view(np.uint64).copy())

    def scan(self):
        allk = np.concatenate(self.recv)
        self.sorted = np.sort(allk, kind="stable")
        d = self.sorted // np.uint64(10)
        self.keep = np.ones(len(d), dtype=bool)
        self.keep[1:] = d[1:] != d[:-1]

    def boundary_bytes(self):
        return 16

    def boundary_get(self):
        return np.array([int(self.sorted[-1]), 1], dtype=np.uint64).tobytes()

    def boundary_fix(self, prev):
        last, valid = np.frombuffer(prev, dtype=np.uint64)
        assert valid == 1
        self.fixed += 1
        self.keep[0] = (self.sorted[0] // np.uint64(10)) != (last // np.uint64(10))

    def emit(self):
        return SimpleNamespace(total=len(self.sorted), dups=int((~self.keep).sum()))

    def output(self, mate):
        return self.sorted[self.keep].tobytes()


def _worker(rank, world, port, slices, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    sh = importlib.import_module("fastq-dupaway_b200.sharded_seq")
    ops = NumpyRangeOps(slices[rank])
    owned, kept, dups = sh.dedup_ranges(ops, dist, rank, world, n_samples=64, tensor_device=torch.device("cpu"))
    out = ops.output(0) if owned else b""
    (Path(result_dir) / f"out_{rank}.bin").write_bytes(out)
    (Path(result_dir) / f"cnt_{rank}.txt").write_text(f"{owned} {kept} {dups} {ops.fixed}")
    dist.barrier()
    dist.destroy_process_group()


def _expected(keys):
    s = np.sort(np.asarray(keys, dtype=np.uint64), kind="stable")
    d = s // np.uint64(10)
    keep = np.ones(len(s), dtype=bool)
    keep[1:] = d[1:] != d[:-1]
    return s[keep]


def _run(tmp_path, slices):
    world = len(slices)
    port = 30900 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, slices, str(tmp_path)), nprocs=world, join=True)
    got = np.frombuffer(b"".join((tmp_path / f"out_{r}.bin").read_bytes() for r in range(world)), dtype=np.uint64)
    cnt = [tuple(int(x) for x in (tmp_path / f"cnt_{r}.txt").read_text().split()) for r in range(world)]
    exp = _expected(np.concatenate([np.asarray(s, dtype=np.uint64) for s in slices]))
    assert np.array_equal(got, exp)
    assert sum(c[0] for c in cnt) == sum(len(s) for s in slices)
    return cnt


def test_three_ranks_dense_keys(tmp_path):
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 3000, size=6000).astype(np.uint64)          # every decade is hit many times
    cnt = _run(tmp_path, [keys[:1000], keys[1000:4500], keys[4500:]])
    assert all(c[0] > 0 for c in cnt)
    assert cnt[1][3] == 1 and cnt[2][3] == 1                            # both later ranges applied a boundary state


def test_empty_slice_and_empty_range(tmp_path):
    # rank 1 contributes nothing; all keys are equal, so two of the three key ranges stay empty and the single owner's
    # boundary state has to pass through them untouched
    keys = np.full(300, 77, dtype=np.uint64)
    cnt = _run(tmp_path, [keys[:100], keys[:0], keys[100:]])
    assert sorted(c[0] for c in cnt) == [0, 0, 300]

"""CPU test (3 gloo ranks) of the multi-GPU --fast --unordered protocol in fastq-dupaway_b200/sharded_unordered.py.
The device side (GpuTagRangeOps) is replaced by a pure-Python stand-in with the same contract that works on real
FASTA / FASTQ bytes, so the concatenated output of the ranks can be compared byte for byte with the oracle's run over
the whole input (which is pinned to the reference's fixtures and binary, tests/test_oracle.py).  Hundreds of tiny cases
per spawn - tag multisets over a small alphabet, random slices, random drops - put the reference's end-of-stream rule
(the walk stops when either side has fetched its last record) on, before and after every range boundary."""
import bisect
import hashlib
import importlib
import os
import pickle
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
NONE = 0xFFFFFFFFFFFFFFFF


def parse_records(buf: bytes, lines_per_record: int):
    lines = buf.split(b"\n")[:-1]
    return [b"".join(x + b"\n" for x in lines[k: k + lines_per_record]) for k in range(0, len(lines), lines_per_record)]


def tag_of(rec: bytes) -> bytes:
    """src/fastqview.cpp:190-204: after the first '.' of the ID line (else after the lead character), up to the first ' '
    at / after the tag start, else to the end of the line including the newline."""
    idline = rec[: rec.index(b"\n") + 1]
    dot = idline.find(b".")
    t0 = dot + 1 if dot >= 0 else 1
    sp = idline.find(b" ", t0)
    return idline[t0: sp if sp >= 0 else len(idline)]


def tag_words(tag: bytes):
    p = tag[:16].ljust(16, b"\0")
    return int.from_bytes(p[:8], "big"), int.from_bytes(p[8:], "big")


class PyTagRangeOps:
    row_bytes = 16

    def __init__(self, slices, lines_per_record):
        self.lpr = lines_per_record
        self.origin = [parse_records(b, lines_per_record) for b in slices]
        self.lists = [[], []]

    # -- origin side
    def sample(self, n):
        out = np.full((n, 2), NONE, dtype=np.uint64)
        half = n // 2
        for m in range(2):
            recs = self.origin[m]
            for j in range(half if recs else 0):
                out[m * half + j] = tag_words(tag_of(recs[(j * len(recs)) // half]))
        return out, len(self.origin[0]) + len(self.origin[1]), 0

    def plan(self, splitters, world):
        sp = [(int(a), int(b)) for a, b in splitters]
        counts = [0] * world
        nbytes = [[0] * world for _ in range(2)]
        self.grouped = []
        for m in range(2):
            owner = [sum(1 for s in sp if s <= tag_words(tag_of(r))) for r in self.origin[m]]
            parts = [[r for r, o in zip(self.origin[m], owner) if o == k] for k in range(world)]
            for k in range(world):
                counts[k] += len(parts[k])
                nbytes[m][k] = sum(len(r) for r in parts[k])
            self.grouped.append(b"".join(b"".join(p) for p in parts))
        return counts, nbytes

    def gather(self, mate, total):
        assert len(self.grouped[mate]) == total
        return torch.frombuffer(bytearray(self.grouped[mate]), dtype=torch.uint8) if total else torch.empty(0, dtype=torch.uint8)

    # -- range side
    def receive(self, mate, recv):
        self.lists[mate] = parse_records(recv.numpy().tobytes(), self.lpr)

    def prepare(self):
        self.S = [sorted(l, key=tag_of) for l in self.lists]          # stable; bytes order = strncmp + shorter-first
        self.T = [[tag_of(r) for r in l] for l in self.S]
        return len(self.S[0]), len(self.S[1]), 0

    def enter(self, side, i):
        A, B = self.T[side], self.T[1 - side]
        if i == 0:
            return 0
        key = A[i - 1]
        rank = (i - 1) - bisect.bisect_left(A, key)
        lb, ub = bisect.bisect_left(B, key), bisect.bisect_right(B, key)
        return lb + min(rank + 1, ub - lb)

    def join(self, lim_i, lim_j, fin_i, fin_j):
        L, R = self.T
        partner = []
        matched_r = set()
        for i, key in enumerate(L):
            rank = i - bisect.bisect_left(L, key)
            lb, ub = bisect.bisect_left(R, key), bisect.bisect_right(R, key)
            pr = lb + rank if rank < ub - lb else None
            partner.append(pr)
            if pr is not None:
                matched_r.add(pr)
        fe = fin_i != NONE and fin_j != NONE and L[fin_i] == R[fin_j]
        self.pairs = []
        for i, pr in enumerate(partner):
            if pr is not None and i < lim_i and pr < lim_j:
                self.pairs.append((i, pr))
            elif fe and i == fin_i:
                self.pairs.append((i, fin_j))
        unmatched = sum(1 for i in range(min(lim_i, len(L))) if partner[i] is None) + \
            sum(1 for j in range(min(lim_j, len(R))) if j not in matched_r)
        bad = None
        for e, (i, j) in enumerate(self.pairs):
            if any(set(self._seq(rec)) - set(b"ACGTN") for rec in (self.S[0][i], self.S[1][j])):
                bad = e
                break
        return len(self.pairs), unmatched, bad, fe

    @staticmethod
    def _seq(rec):
        return rec.split(b"\n")[1]

    def rows(self, limit, world):
        digests = [hashlib.md5(self._seq(self.S[0][i]) + b"|" + self._seq(self.S[1][j])).digest() for i, j in self.pairs[:limit]]
        owner = [(int.from_bytes(d[:8], "big") * world) >> 64 for d in digests]
        self.send_idx = sorted(range(limit), key=lambda e: owner[e])          # stable
        counts = [owner.count(k) for k in range(world)]
        data = b"".join(digests[e] for e in self.send_idx)
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8).reshape(-1, 16) if limit else torch.empty((0, 16), dtype=torch.uint8)
        return t, counts

    def insert(self, recv_rows, world):
        seen = set()
        flags = []
        for row in recv_rows.numpy().reshape(-1, 16):
            k = row.tobytes()
            flags.append(1 if k in seen else 0)
            seen.add(k)
        return torch.tensor(flags, dtype=torch.uint8)

    def apply(self, flags_back, limit, report_bad):
        dup = [0] * len(self.pairs)
        for pos, e in enumerate(self.send_idx):
            dup[e] = int(flags_back[pos])
        self.kept = [p for e, p in enumerate(self.pairs) if e < limit and not dup[e]]
        self.reported_bad = report_bad
        return SimpleNamespace(total=limit, dups=limit - len(self.kept))

    def output(self, mate):
        return b"".join(self.S[mate][p[mate]] for p in self.kept)


def _worker(rank, world, port, case_file, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("fastq-dupaway_b200.sharded_unordered")
    cases = pickle.loads(Path(case_file).read_bytes())
    results = []
    for lpr, slices, n_samples in cases:
        ops = PyTagRangeOps(slices[rank], lpr)
        res = sh.dedup_tag_ranges(ops, dist, rank, world, n_samples=n_samples, tensor_device=torch.device("cpu"))
        outs = (b"", b"") if res.err == sh.ERR_EMPTY else (ops.output(0), ops.output(1))
        results.append((res, outs))
    (Path(result_dir) / f"res_{rank}.pkl").write_bytes(pickle.dumps(results))
    dist.barrier()
    dist.destroy_process_group()


def run_cases(tmp_path, cases, world, worker=_worker, port_base=31900):
    """cases: [(lines per record, slices[rank] = (bytes of file 1, bytes of file 2), n_samples)] -> per case
    (JobResult of rank 0, [per-rank JobResult], output 1, output 2)"""
    case_file = tmp_path / "cases.pkl"
    case_file.write_bytes(pickle.dumps(cases))
    port = port_base + (os.getpid() % 2000)
    mp.spawn(worker, args=(world, port, str(case_file), str(tmp_path)), nprocs=world, join=True)
    per_rank = [pickle.loads((tmp_path / f"res_{r}.pkl").read_bytes()) for r in range(world)]
    out = []
    for c in range(len(cases)):
        rs = [per_rank[r][c][0] for r in range(world)]
        out.append((rs[0], rs, b"".join(per_rank[r][c][1][0] for r in range(world)), b"".join(per_rank[r][c][1][1] for r in range(world))))
    return out


ERRMAP = {0: 0, 1: 3, 2: 4, 3: 5, 4: 6}


def check_against_oracle(oracle, fmt, whole, got):
    res, per_rank, o1, o2 = got
    e1, e2, est = oracle.run_oracle("fast", fmt, whole[0], whole[1], unordered=True)
    assert all((r.err, r.total, r.dups, r.unmatched) == (res.err, res.total, res.dups, res.unmatched) for r in per_rank)
    assert res.err == ERRMAP[est.err], (res, est.err)
    if est.err == 1:
        return
    assert (o1, o2) == (e1, e2), (whole, o1, e1)
    assert (res.total, res.dups) == (est.total, est.dups), (whole, res, est.total, est.dups)
    if est.err == 0:            # an aborted run has not counted the skipped entries behind the abort point
        assert res.unmatched == est.unmatched, (whole, res, est.unmatched)
    assert sum(r.local_kept for r in per_rank) == res.total - res.dups


def slice_records(recs, world, rng):
    cuts = sorted(int(x) for x in rng.integers(0, len(recs) + 1, size=world - 1))
    bounds = [0] + cuts + [len(recs)]
    return [b"".join(recs[bounds[r]: bounds[r + 1]]) for r in range(world)]


def tiny_cases(n_cases, seed, world, max_len=6, n_tags=5):
    rng = np.random.default_rng(seed)
    seqs = [b"ACGT", b"ACGA", b"TTTT"]
    cases, wholes = [], []
    for _ in range(n_cases):
        files = []
        for mate in range(2):
            k = int(rng.integers(1, max_len + 1))
            recs = [b">X.%d %d\n%s\n" % (int(rng.integers(1, n_tags + 1)), mate + 1, seqs[int(rng.integers(0, 3))]) for _ in range(k)]
            files.append(recs)
        sl = [slice_records(f, world, rng) for f in files]
        cases.append((2, [(sl[0][r], sl[1][r]) for r in range(world)], int(rng.choice([2, 4, 8]))))
        wholes.append((b"".join(files[0]), b"".join(files[1])))
    return cases, wholes


def test_stop_rule_over_every_boundary(tmp_path, oracle):
    cases, wholes = tiny_cases(400, seed=11, world=3)
    got = run_cases(tmp_path, cases, 3)
    for whole, g in zip(wholes, got):
        check_against_oracle(oracle, oracle.FASTA, whole, g)


def big_cases(seed, world=3, n=600):
    """shuffled FASTQ files with deletions and duplicate pairs; variant 2: a byte outside {A,C,G,T,N} in a matched pair
    (the job aborts there); variant 3: one tag many times on both sides (pairs are matched by rank inside the tag)"""
    import synth
    rng = np.random.default_rng(seed)
    cases, wholes = [], []
    for variant in range(4):
        s1, s2 = synth.make_pair(n, seed=seed + 8 + variant, read_len=24, dup_frac=0.4)
        ids = [b"@RUN.%05d" % i for i in range(n)]
        a = [synth.to_fastq([s1[i]], ids=[ids[i] + b" 1"]) for i in range(n) if rng.random() > 0.1]
        b = [synth.to_fastq([s2[i]], ids=[ids[i] + b" 2"]) for i in range(n) if rng.random() > 0.1]
        if variant == 2:
            k = len(a) // 2
            lines = a[k].split(b"\n")
            lines[1] = b"X" + lines[1][1:]
            a[k] = b"\n".join(lines)
        if variant == 3:
            a = [r.replace(r[5:10], b"00007", 1) if i % 3 == 0 else r for i, r in enumerate(a)]
            b = [r.replace(r[5:10], b"00007", 1) if i % 4 == 0 else r for i, r in enumerate(b)]
        order_a, order_b = rng.permutation(len(a)), rng.permutation(len(b))
        a, b = [a[int(k)] for k in order_a], [b[int(k)] for k in order_b]
        sl = [slice_records(a, world, rng), slice_records(b, world, rng)]
        cases.append((4, [(sl[0][r], sl[1][r]) for r in range(world)], 64))
        wholes.append((b"".join(a), b"".join(b)))
    return cases, wholes


def test_shuffled_files_with_deletions_duplicates_and_a_bad_base(tmp_path, oracle):
    cases, wholes = big_cases(12)
    got = run_cases(tmp_path, cases, 3)
    for whole, g in zip(wholes, got):
        check_against_oracle(oracle, oracle.FASTQ, whole, g)
    assert got[2][0].err == 6


def test_one_file_empty(tmp_path, oracle):
    cases = [(2, [(b">X.1\nACGT\n", b""), (b"", b""), (b">X.2\nACGT\n", b"")], 4)]
    got = run_cases(tmp_path, cases, 3)
    assert got[0][0].err == 3


def test_one_rank_is_the_whole_job(tmp_path, oracle):
    cases, wholes = tiny_cases(60, seed=13, world=1)
    got = run_cases(tmp_path, cases, 1)
    for whole, g in zip(wholes, got):
        check_against_oracle(oracle, oracle.FASTA, whole, g)

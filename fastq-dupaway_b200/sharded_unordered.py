"""Multi-GPU --fast --unordered: ID-TAG-range sharding of both files across the ranks of one box (the
reference's single-threaded counterpart is src/hash_dup_remover.hpp:150-192,257-347).

One process per GPU; rank r holds a contiguous slice of file 1 and a contiguous slice of file 2 (any slices).
  1. every rank parses its slices (origin handle) and samples the tags of both files         fqd_partition_sample
  2. all-gather of the samples -> world-1 tag splitters, the same on every rank
  3. every file is partitioned by ITS OWN tags; one all-to-all per file moves the raw records to the rank that owns
     their tag range (source-rank major = input order, so "k-th record of a tag" survives)   fqd_partition_plan / _gather
  4. every rank sorts its two tag lists and finds the partner of every record               fqd_unordered_prepare
  5. the walk of the reference stops when either side has fetched its LAST record.  With the list
     lengths of all ranges known (all-gather), the rank holding L[n-2] and the rank holding R[m-2] say where the walk
     stands at that moment (fqd_unordered_enter); every rank derives the job's stop state (is, js) and its own share
  6. emitted pairs per range, their keys, the first pair with a byte outside {A,C,G,T,N}     fqd_unordered_join
  7. all-gather of the counts: emission offsets (ranges in rank order = tag order = the reference's emission order),
     the job's abort point, the unmatched total
  8. first-occurrence dedup over the whole job: pair keys go to the owner of their HASH range with one all-to-all
     (what arrives is in emission order), the owner flags every key it has seen before, the flags travel back
                                                                                     fqd_unordered_rows / _insert / _apply
The job's output is the concatenation of the ranks' outputs in rank order.
The plumbing is independent of the device code (`ops` is any object with the methods of GpuTagRangeOps), which is how
tests/test_sharded_unordered_cpu.py runs it with three gloo ranks and a pure-Python stand-in.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .sharded_seq import choose_splitters

NONE = 0xFFFFFFFFFFFFFFFF
ERR_EMPTY = 3
ERR_BAD_BASE = 6


@dataclass
class JobResult:
    err: int            # 0, ERR_EMPTY (one of the files holds no record) or ERR_BAD_BASE (the job aborted at `total`)
    total: int          # valid read pairs processed (whole job)
    dups: int           # duplicates among them (whole job)
    unmatched: int      # non-matching entries skipped (whole job)
    local_pairs: int    # pairs this rank emitted before deduplication
    local_kept: int     # pairs this rank writes


class GpuTagRangeOps:
    """The device side of one rank: an origin handle for this rank's slices, a range handle for the tags it owns."""

    def __init__(self, pkg, fmt, max_seq_len, max_records_origin, max_records_range, device, seg_bytes=1 << 28, max_tag_len=0):
        import torch
        self.torch = torch
        self.pkg = pkg
        self.dev = torch.device("cuda", device)
        self.origin = pkg.Engine("fast", fmt, True, True, 2, max_seq_len, max_records_origin, seg_bytes, 0, device, max_tag_len)
        self.range = pkg.Engine("fast", fmt, True, True, 2, max_seq_len, max_records_range, seg_bytes, 0, device, max_tag_len)
        self.row_bytes = int(self.range.lib.fqd_unordered_row_bytes(self.range.h))
        self._keep = []         # device tensors the handles still read

    # -- origin side
    def append(self, mate, dptr, nbytes):
        self.origin.append_device(mate, dptr, nbytes)

    def adopt(self, mate, dptr, nbytes):
        """this rank's whole slice of one file, device-resident and 16-byte aligned: parsed in place, not copied"""
        self.origin.adopt_device(mate, dptr, nbytes)

    def sample(self, n_samples):
        smp, n = self.origin.partition_sample(n_samples)
        return smp, n, int(self.origin.stats().err)

    def plan(self, splitters, world):
        return self.origin.partition_plan(splitters, world)

    def gather(self, mate, total_bytes):
        buf = self.torch.empty(max(1, total_bytes), dtype=self.torch.uint8, device=self.dev)
        self.origin.partition_gather(mate, buf.data_ptr())
        self.torch.cuda.synchronize(self.dev)
        return buf[:total_bytes]

    # -- range side
    def receive(self, mate, recv):
        """what the all-to-all delivered is parsed in place (the tensor stays alive until reset)"""
        self.torch.cuda.synchronize(self.dev)
        if recv.numel():
            self._keep.append(recv)
            self.range.adopt_device(mate, recv.data_ptr(), int(recv.numel()))

    def prepare(self):
        nl, nr = self.range.unordered_prepare()
        return nl, nr, int(self.range.stats().err)

    def enter(self, side, i):
        return self.range.unordered_enter(side, i)

    def join(self, limit_i, limit_j, final_i, final_j):
        return self.range.unordered_join(limit_i, limit_j, final_i, final_j)

    def rows(self, limit, world):
        send = self.torch.empty((max(1, limit), self.row_bytes), dtype=self.torch.uint8, device=self.dev)
        counts = self.range.unordered_rows(limit, world, send.data_ptr())
        self.torch.cuda.synchronize(self.dev)
        return send[:limit], counts

    def insert(self, recv_rows, world):
        n = int(recv_rows.shape[0])
        flags = self.torch.zeros(max(1, n), dtype=self.torch.uint8, device=self.dev)
        self.torch.cuda.synchronize(self.dev)
        if n:
            self.range.unordered_insert(recv_rows.data_ptr(), n, world, flags.data_ptr())
        return flags[:n]

    def apply(self, flags_back, limit, report_bad):
        self.torch.cuda.synchronize(self.dev)
        self._keep.append(flags_back)
        self.range.unordered_apply(flags_back.data_ptr() if flags_back.numel() else 0, limit, report_bad)
        return self.range.stats()

    def output(self, mate):
        return self.range.emit_all(mate)

    def device_ms(self):
        return self.origin.device_time_ms()[0] + self.range.device_time_ms()[0]

    def reset(self):
        self.origin.reset()
        self.range.reset()
        self.torch.cuda.synchronize(self.dev)
        self._keep.clear()

    def close(self):
        self.origin.close()
        self.range.close()


def _gather_ints(dist, values, world, dev):
    """all-gather of a short list of integers (uint64 values travel as int64 bit patterns) -> [world][len(values)]"""
    import torch
    mine = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in values], dtype=torch.int64, device=dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [[int(x) & NONE for x in t.cpu().tolist()] for t in out]


def _all_to_all_bytes(dist, send, send_sizes, recv_sizes, via_cpu):
    import torch
    if via_cpu:
        r_h = torch.empty(sum(recv_sizes), dtype=torch.uint8)
        dist.all_to_all_single(r_h, send.reshape(-1).cpu(), recv_sizes, send_sizes)
        return r_h.to(send.device)
    recv = torch.empty(sum(recv_sizes), dtype=torch.uint8, device=send.device)
    dist.all_to_all_single(recv, send.reshape(-1), recv_sizes, send_sizes)
    return recv


def _exchange_sizes(dist, sizes, world, dev):
    import torch
    c_out = torch.tensor(sizes, dtype=torch.int64, device=dev)
    c_in = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(c_in, c_out)
    return [int(x) for x in c_in.tolist()]


def stop_state(n, m, ja, ia):
    """The walk `while i < n-1 and j < m-1` (src/hash_dup_remover.hpp:279-281) ends in (is, js): ja = R position when it
    first stands on L[n-1], ia = L position when it first stands on R[m-1] (both as if the other side never ended)."""
    return (n - 1, ja) if ja < m - 1 else (ia, m - 1)


def dedup_tag_ranges(ops, dist, rank, world, n_samples=4096, tensor_device=None, via_cpu=False):
    """Steps 1-8 above for what has been appended to ops' origin handle.  Returns a JobResult; `ops.output(mate)` then
    yields this rank's part of the output (empty when err == ERR_EMPTY)."""
    import torch
    dev = tensor_device if tensor_device is not None else getattr(ops, "dev", torch.device("cpu"))
    cdev = torch.device("cpu") if via_cpu else dev          # gloo ranks sharing one GPU (tests): collectives on host tensors
    # 1-2. samples -> splitters; a data error anywhere stops every rank
    samples, n_local, err = ops.sample(n_samples)
    errs = _gather_ints(dist, [err], world, cdev)
    if any(e[0] for e in errs):
        bad = next(r for r, e in enumerate(errs) if e[0])
        raise RuntimeError(f"rank {bad}: data error {errs[bad][0]} in its slice of the input")
    mine = torch.from_numpy(samples.astype(np.int64)).to(cdev)
    allsmp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allsmp, mine)
    splitters = choose_splitters(np.concatenate([t.cpu().numpy().astype(np.uint64) for t in allsmp], axis=0), world)
    # 3. every file to the owners of its tags
    counts, nbytes = ops.plan(splitters, world)
    if sum(counts) != n_local:
        raise RuntimeError(f"rank {rank}: partition plan covers {sum(counts)} of {n_local} records")
    for m in range(2):
        send = ops.gather(m, sum(nbytes[m]))
        recv_sizes = _exchange_sizes(dist, nbytes[m], world, cdev)
        ops.receive(m, _all_to_all_bytes(dist, send, nbytes[m], recv_sizes, via_cpu))
    # 4. local tag sort + partners
    nl, nr, err = ops.prepare()
    lens = _gather_ints(dist, [nl, nr, err], world, cdev)
    if any(e[2] for e in lens):
        bad = next(r for r, e in enumerate(lens) if e[2])
        raise RuntimeError(f"rank {bad}: data error {lens[bad][2]} in its tag range")
    NL, NR = [e[0] for e in lens], [e[1] for e in lens]
    offL, offR = np.concatenate([[0], np.cumsum(NL)]).tolist(), np.concatenate([[0], np.cumsum(NR)]).tolist()
    n, m = offL[-1], offR[-1]
    if n == 0 or m == 0:
        return JobResult(ERR_EMPTY, 0, 0, 0, 0, 0)

    # 5. where the walk stops
    def holder(N, off, g):
        return next(p for p in range(world) if off[p] <= g < off[p] + N[p])
    ja = ia = NONE
    if n == 1:
        ja = 0
    elif holder(NL, offL, n - 2) == rank:
        ja = offR[rank] + ops.enter(0, n - 2 - offL[rank] + 1)
    if m == 1:
        ia = 0
    elif holder(NR, offR, m - 2) == rank:
        ia = offL[rank] + ops.enter(1, m - 2 - offR[rank] + 1)
    ent = _gather_ints(dist, [ja, ia], world, cdev)
    ja = next(e[0] for e in ent if e[0] != NONE)
    ia = next(e[1] for e in ent if e[1] != NONE)
    i_s, j_s = stop_state(n, m, ja, ia)
    lim_i = min(max(i_s - offL[rank], 0), nl)
    lim_j = min(max(j_s - offR[rank], 0), nr)
    fin_i = i_s - offL[rank] if offL[rank] <= i_s < offL[rank] + nl else NONE
    fin_j = j_s - offR[rank] if offR[rank] <= j_s < offR[rank] + nr else NONE
    # 6-7. emitted pairs, job-wide bookkeeping
    E, unmatched, bad, fe = ops.join(lim_i, lim_j, fin_i, fin_j)
    info = _gather_ints(dist, [E, unmatched, NONE if bad is None else bad, int(fe)], world, cdev)
    offE = np.concatenate([[0], np.cumsum([e[0] for e in info])]).tolist()
    unmatched_job = sum(e[1] for e in info) + (0 if any(e[3] for e in info) else 1)
    bad_at = [(offE[p] + e[2], p) for p, e in enumerate(info) if e[2] != NONE]
    abort, bad_rank = min(bad_at) if bad_at else (None, None)          # the run aborts when that pair is keyed
    limit = E if abort is None else min(max(abort - offE[rank], 0), E)
    # 8. first occurrence over the whole job
    rows, rcounts = ops.rows(limit, world)
    rb = ops.row_bytes
    recv_counts = _exchange_sizes(dist, rcounts, world, cdev)
    recv = _all_to_all_bytes(dist, rows, [c * rb for c in rcounts], [c * rb for c in recv_counts], via_cpu)
    flags = ops.insert(recv.reshape(-1, rb), world)
    back = _all_to_all_bytes(dist, flags, recv_counts, rcounts, via_cpu)
    st = ops.apply(back, limit, bad_rank == rank)
    kept = int(st.total - st.dups)
    tot = _gather_ints(dist, [limit, kept], world, cdev)
    total_job = sum(t[0] for t in tot)
    return JobResult(ERR_BAD_BASE if abort is not None else 0, total_job, total_job - sum(t[1] for t in tot), unmatched_job, E, kept)

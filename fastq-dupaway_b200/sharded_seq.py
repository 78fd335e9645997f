"""Multi-GPU sequence-based mode: key-RANGE sharding of the sorted stream across the ranks of one box.

One process per GPU; rank r holds a contiguous slice of the input (slices in rank order = the input).
  1. every rank parses + packs its slice (origin engine)                                      fqd_partition_sample
  2. all-gather of sampled (word 0, word 1) key pairs -> world-1 splitters, the same on every rank
  3. owner of every record, records grouped by owner                                          fqd_partition_plan / _gather
  4. ONE all-to-all per mate moves the raw records to the rank that owns their key range (NCCL over NVLink);
     what arrives is in global input order (source rank major, input order inside a source), so the stable tie-break
     "first in input order" survives the exchange
  5. every rank runs the ordinary engine on what it received: sort + comparator scan        fqd_finish_scan
  6. boundary states travel rank k-1 -> rank k (a few hundred bytes each): the first sorted records of a range are
     re-evaluated against the last record / last cluster head of the range before             fqd_boundary_get / _fix
  7. emission lists                                                                             fqd_finish_emit
The job's output is the concatenation of the ranks' outputs in rank order (the reference writes sorted order).
The plumbing is independent of the device code (`ops` is any object with the methods of GpuRangeOps), which is how
tests/test_sharded_seq_cpu.py exercises it with two gloo ranks and a CPU stand-in.
"""
from __future__ import annotations

import numpy as np


def choose_splitters(samples: np.ndarray, world: int) -> np.ndarray:
    """samples: uint64 [k, 2] (all-ones rows = no data) -> [world - 1, 2] ascending splitters (quantiles)."""
    if world <= 1:
        return np.zeros((0, 2), dtype=np.uint64)
    s = samples[~((samples[:, 0] == np.uint64(0xFFFFFFFFFFFFFFFF)) & (samples[:, 1] == np.uint64(0xFFFFFFFFFFFFFFFF)))]
    if len(s) == 0:
        return np.full((world - 1, 2), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    order = np.lexsort((s[:, 1], s[:, 0]))
    s = s[order]
    pick = [(len(s) * (k + 1)) // world for k in range(world - 1)]
    return s[np.minimum(pick, len(s) - 1)]


class GpuRangeOps:
    """The device side of one rank: an origin engine for this rank's slice and a range engine for what it owns."""

    def __init__(self, pkg, mode, fmt, paired, dist, max_seq_len, max_records_origin, max_records_range, device, seg_bytes=1 << 28):
        import torch
        self.torch = torch
        self.pkg = pkg
        self.paired = paired
        self.mates = 2 if paired else 1
        self.dev = torch.device("cuda", device)
        self.origin = pkg.Engine(mode, fmt, paired, False, dist, max_seq_len, max_records_origin, seg_bytes, 0, device)
        self.range = pkg.Engine(mode, fmt, paired, False, dist, max_seq_len, max_records_range, seg_bytes, 0, device)

    # -- origin side
    def append(self, mate, dptr, nbytes):
        self.origin.append_device(mate, dptr, nbytes)

    def sample(self, n_samples):
        return self.origin.partition_sample(n_samples)

    def plan(self, splitters, world):
        return self.origin.partition_plan(splitters, world)

    def gather(self, mate, total_bytes):
        buf = self.torch.empty(max(1, total_bytes), dtype=self.torch.uint8, device=self.dev)
        self.origin.partition_gather(mate, buf.data_ptr())
        self.torch.cuda.synchronize(self.dev)
        return buf[:total_bytes]

    # -- owner side
    def receive(self, mate, recv):
        self.torch.cuda.synchronize(self.dev)          # the all-to-all that fills recv runs on torch's stream
        if recv.numel():
            self.range.append_device(mate, recv.data_ptr(), int(recv.numel()))
            self.torch.cuda.synchronize(self.dev)

    def receive_ptr(self, mate, dptr, nbytes):
        """what arrived in a peer-exchange buffer is parsed in place (the buffer is not reused before the job ends)"""
        if nbytes:
            self.range.adopt_device(mate, dptr, nbytes)

    def scan(self):
        self.range.finish_scan()
        st = self.range.stats()
        if st.err:
            raise self.pkg.FqdError(st.err, f"key range engine: data error at record {st.err_record}")

    def boundary_bytes(self):
        return int(self.range.lib.fqd_boundary_bytes(self.range.h))

    def boundary_get(self):
        return self.range.boundary_get()

    def boundary_fix(self, prev):
        self.range.boundary_fix(prev)

    def emit(self):
        self.range.finish_emit()
        return self.range.stats()

    def output(self, mate):
        return self.range.emit_all(mate)

    def reset(self):
        self.origin.reset()
        self.range.reset()

    def close(self):
        self.origin.close()
        self.range.close()


TRACE = {}


def _mark(name, t0):
    """FQD_TRACE=1: wall clock per phase of dedup_ranges (each mark synchronises the device)."""
    import os
    import time
    if not os.environ.get("FQD_TRACE"):
        return t0
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    TRACE[name] = TRACE.get(name, 0.0) + (t1 - t0) * 1e3
    return t1


def dedup_ranges(ops, dist, rank, world, n_samples=4096, tensor_device=None, via_cpu=False, peer=None):
    """Steps 1-7 above for what has been appended to ops' origin engine.  peer: one PeerExchange per mate - the records
    then travel over mapped peer memory and are parsed in place where they land; else torch.distributed moves them.  Returns (records this rank owns,
    records it writes, duplicates it removed) - `ops.output(mate)` then yields this rank's part of the output."""
    import torch
    dev = tensor_device if tensor_device is not None else getattr(ops, "dev", torch.device("cpu"))
    if via_cpu:                       # gloo ranks sharing one GPU (tests): collectives on host tensors
        dev = torch.device("cpu")
    import time
    t = time.perf_counter()
    samples, n_local = ops.sample(n_samples)
    t = _mark("sample (parse rest + samples)", t)
    # 2. splitters
    mine = torch.from_numpy(samples.astype(np.int64)).to(dev)
    allsmp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allsmp, mine)
    gathered = np.concatenate([t.cpu().numpy().astype(np.uint64) for t in allsmp], axis=0)
    splitters = choose_splitters(gathered, world)
    t = _mark("splitters (all-gather + sort)", t)
    # 3. plan
    counts, nbytes = ops.plan(splitters, world)
    t = _mark("plan (owner + partition + offsets)", t)
    mates = len(nbytes)
    if sum(counts) != n_local:
        raise RuntimeError(f"rank {rank}: partition plan covers {sum(counts)} of {n_local} records")
    # 4. exchange
    c_out = torch.tensor([counts] + nbytes, dtype=torch.int64, device=dev).reshape(-1)
    c_in = torch.empty_like(c_out)
    dist.all_to_all_single(c_in, c_out.reshape(1 + mates, world).t().contiguous().reshape(-1))
    c_in = c_in.reshape(world, 1 + mates).t().contiguous()         # [1 + mates][source]
    n_owned = int(c_in[0].sum().item())
    t = _mark("counts all-to-all", t)
    for m in range(mates):
        send = ops.gather(m, sum(nbytes[m]))
        t = _mark("gather records", t)
        if peer is not None:          # mapped peer memory: the copy engines write straight into the owners' buffers
            rptr, rsizes = peer[m].exchange(send.data_ptr(), nbytes[m])
            t = _mark("records all-to-all", t)
            ops.receive_ptr(m, rptr, sum(rsizes))
            t = _mark("receive (append + parse of full segments)", t)
            continue
        recv_sizes = [int(x) for x in c_in[1 + m].tolist()]
        if via_cpu:
            r_h = torch.empty(sum(recv_sizes), dtype=torch.uint8)
            dist.all_to_all_single(r_h, send.cpu(), recv_sizes, nbytes[m])
            recv = r_h.to(send.device)
        else:
            recv = torch.empty(sum(recv_sizes), dtype=torch.uint8, device=send.device)
            dist.all_to_all_single(recv, send, recv_sizes, nbytes[m])
        t = _mark("records all-to-all", t)
        ops.receive(m, recv)
        t = _mark("receive (append + parse of full segments)", t)
    # 5. local sort + scan
    have = n_owned > 0
    if have:
        ops.scan()
    t = _mark("scan stage (parse rest + sort + scan)", t)
    # 6. boundary chain: the state after range k-1 goes to range k; an empty range passes on what it received
    nb = ops.boundary_bytes()
    state = torch.zeros(nb, dtype=torch.uint8, device=dev)          # all-zero = "nothing before" (valid flag 0)
    for k in range(1, world):
        if rank == k - 1:
            if have:
                state = torch.frombuffer(bytearray(ops.boundary_get()), dtype=torch.uint8).to(dev)
        dist.broadcast(state, src=k - 1)
        if rank == k and have:
            prev = bytes(state.cpu().numpy().tobytes())
            if any(prev):
                ops.boundary_fix(prev)
        # rank k now holds, in `state`, what precedes it; if it has records it will overwrite it with its own tail
    t = _mark("boundary chain", t)
    # 7. emission
    if have:
        st = ops.emit()
        return n_owned, int(st.total - st.dups), int(st.dups)
    return 0, 0, 0

"""Multi-GPU --fast mode: hash-range sharding of the key space across the ranks of one box.

One process per GPU.  Per chunk every rank
  1. splits + packs its own slice of the input and groups the packed keys by owning rank   (fqd_shard_pack)
  2. exchanges the key rows with ONE all-to-all (torch.distributed: NCCL over NVLink / NVSwitch)
  3. inserts the rows it received - they arrive ordered by global input position - into its set (fqd_shard_insert)
  4. returns one duplicate flag per row with a second, byte-sized all-to-all
  5. stores the flags against its own records                                                (fqd_shard_apply)
The exchange plumbing below is independent of the device code (`ops` is any object with pack / insert / apply), which
is how tests/test_sharded_cpu.py exercises it with two gloo ranks on the CPU.
"""
from __future__ import annotations

import ctypes as C


class GpuShardOps:
    """pack / insert / apply on one GPU through the C ABI; buffers are torch tensors on that GPU."""

    def __init__(self, pkg, engine, world, device, max_rows, own_stream=False):
        import torch
        self.torch = torch
        self.pkg, self.eng, self.world = pkg, engine, world
        self.lib = pkg.load_library()
        self.dev = torch.device("cuda", device)
        self.row_bytes = int(self.lib.fqd_shard_row_bytes(engine.h))
        self.send = torch.empty((max_rows, self.row_bytes), dtype=torch.uint8, device=self.dev)
        self.flags = torch.empty(2 * max_rows, dtype=torch.uint8, device=self.dev)
        # own_stream: a packer that works ahead of the exchange (exchange_pipelined) must not queue behind it
        self.stream = torch.cuda.Stream(device=self.dev) if own_stream else torch.cuda.current_stream(self.dev)
        rc = self.lib.fqd_set_stream(engine.h, C.c_void_p(self.stream.cuda_stream))
        assert rc == 0

    def pack(self, raw_ptr, nbytes, raw2_ptr=None, nbytes2=0):
        counts = (C.c_uint64 * self.world)()
        nrec = C.c_uint64(0)
        if raw2_ptr is None:
            rc = self.lib.fqd_shard_pack(self.eng.h, C.c_void_p(raw_ptr), nbytes, self.world, C.c_void_p(self.send.data_ptr()), counts, C.byref(nrec))
        else:
            rc = self.lib.fqd_shard_pack_pe(self.eng.h, C.c_void_p(raw_ptr), nbytes, C.c_void_p(raw2_ptr), nbytes2, self.world,
                                            C.c_void_p(self.send.data_ptr()), counts, C.byref(nrec))
        self.eng._check(rc)
        return self.send[: nrec.value], [int(c) for c in counts]

    def insert(self, recv_rows):
        n = int(recv_rows.shape[0])
        rc = self.lib.fqd_shard_insert(self.eng.h, C.c_void_p(recv_rows.data_ptr()), n, self.world, C.c_void_p(self.flags.data_ptr()))
        self.eng._check(rc)
        return self.flags[:n]

    def insert_ptr(self, recv_ptr, n):
        rc = self.lib.fqd_shard_insert(self.eng.h, C.c_void_p(recv_ptr), n, self.world, C.c_void_p(self.flags.data_ptr()))
        self.eng._check(rc)
        return self.flags[:n]

    def read_flags(self, n):
        """duplicate flags of the chunk last applied on this engine (host copy)"""
        out = C.create_string_buffer(int(n))
        assert self.lib.fqd_shard_read_flags(self.eng.h, out, int(n)) == 0
        return out.raw

    def apply(self, flags_back):
        d = C.c_uint64(0)
        rc = self.lib.fqd_shard_apply(self.eng.h, C.c_void_p(flags_back.data_ptr()), C.byref(d))
        self.eng._check(rc)
        return int(d.value)


TRACE = {}


def _mark(name, t0):
    """FQD_TRACE=1: wall-clock per phase of exchange_chunk (each mark synchronises the device)."""
    import os
    import time
    import torch
    if not os.environ.get("FQD_TRACE"):
        return t0
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    TRACE[name] = TRACE.get(name, 0.0) + (t1 - t0) * 1e3
    return t1


def exchange_chunk(ops, dist, world, raw_ptr, nbytes, via_cpu=False, raw2_ptr=None, nbytes2=0, peer=None):
    """One chunk through pack -> all-to-all -> insert -> all-to-all -> apply.  Returns this rank's duplicate count."""
    import time
    import torch
    t = time.perf_counter()
    send_rows, counts = ops.pack(raw_ptr, nbytes) if raw2_ptr is None else ops.pack(raw_ptr, nbytes, raw2_ptr, nbytes2)
    t = _mark("pack (K1 + owner sort + row gather)", t)
    dev = send_rows.device
    row_bytes = int(send_rows.shape[1])
    if peer is not None:
        # rows over mapped peer memory (fastq-dupaway_b200/peer.py): the size matrix replaces the counts all-to-all
        rptr, rsizes = peer.exchange(send_rows.data_ptr(), [c * row_bytes for c in counts])
        recv_counts = [b // row_bytes for b in rsizes]
        n_recv = sum(recv_counts)
        t = _mark("rows all-to-all", t)
        flags = ops.insert_ptr(rptr, n_recv)
        t = _mark("insert (append + K2)", t)
        back = torch.empty(sum(counts), dtype=torch.uint8, device=dev)
        dist.all_to_all_single(back, flags, counts, recv_counts)
        t = _mark("flags all-to-all", t)
        r = ops.apply(back), sum(counts)
        _mark("apply", t)
        return r
    # how many rows will every peer send me?
    cdev = torch.device("cpu") if via_cpu else dev
    c_out = torch.tensor(counts, dtype=torch.int64, device=cdev)
    c_in = torch.empty(world, dtype=torch.int64, device=cdev)
    dist.all_to_all_single(c_in, c_out)
    recv_counts = [int(x) for x in c_in.tolist()]
    n_recv = sum(recv_counts)
    t = _mark("counts all-to-all", t)
    # rows: one all-to-all
    s = send_rows.reshape(-1)
    if via_cpu:
        s_h = s.cpu()
        r_h = torch.empty(n_recv * row_bytes, dtype=torch.uint8)
        dist.all_to_all_single(r_h, s_h, [c * row_bytes for c in recv_counts], [c * row_bytes for c in counts])
        recv = r_h.to(dev)
    else:
        recv = torch.empty(n_recv * row_bytes, dtype=torch.uint8, device=dev)
        dist.all_to_all_single(recv, s, [c * row_bytes for c in recv_counts], [c * row_bytes for c in counts])
    t = _mark("rows all-to-all", t)
    flags = ops.insert(recv.reshape(n_recv, row_bytes))
    t = _mark("insert (append + K2)", t)
    # flags back: one byte per row, reverse direction
    if via_cpu:
        f_h = flags.cpu()
        b_h = torch.empty(sum(counts), dtype=torch.uint8)
        dist.all_to_all_single(b_h, f_h, counts, recv_counts)
        back = b_h.to(dev)
    else:
        back = torch.empty(sum(counts), dtype=torch.uint8, device=dev)
        dist.all_to_all_single(back, flags, counts, recv_counts)
    t = _mark("flags all-to-all", t)
    r = ops.apply(back), sum(counts)
    _mark("apply", t)
    return r


def exchange_pipelined(packers, main, dist, world, chunks, peers, via_cpu=False, flags_out=None):
    """All chunks of one job with the exchange of chunk c overlapping the split + pack of chunk c + 1.
    packers: two GpuShardOps on pack-only engines with their own streams (they alternate); main: the GpuShardOps of the
    engine that owns this rank's hash set; peers: two PeerExchange (alternating receive buffers); chunks: [(ptr, nbytes)].
    The copy engines move chunk c's key rows into the owners' buffers while the SMs split chunk c + 1.  Returns dups."""
    import torch
    n = len(chunks)
    if n == 0:
        return 0
    dev = main.dev
    dups = 0
    send, counts = packers[0].pack(*chunks[0])
    for c in range(n):
        pk, px = packers[c % 2], peers[c % 2]
        rb = pk.row_bytes
        rptr, rsizes = px.start(send.data_ptr(), [k * rb for k in counts])
        cur_counts = counts
        if c + 1 < n:
            send, counts = packers[(c + 1) % 2].pack(*chunks[c + 1])      # K1 of the next chunk runs under the copies
        px.finish()
        recv_counts = [b // rb for b in rsizes]
        flags = main.insert_ptr(rptr, sum(recv_counts))
        if via_cpu:                                   # gloo ranks sharing one GPU (tests)
            b_h = torch.empty(sum(cur_counts), dtype=torch.uint8)
            dist.all_to_all_single(b_h, flags.cpu(), cur_counts, recv_counts)
            back = b_h.to(dev)
        else:
            back = torch.empty(sum(cur_counts), dtype=torch.uint8, device=dev)
            dist.all_to_all_single(back, flags, cur_counts, recv_counts)
        torch.cuda.current_stream(dev).synchronize()
        dups += pk.apply(back)
        if flags_out is not None:                     # per-record duplicate flags of this chunk, in input order
            flags_out.append(pk.read_flags(sum(cur_counts)))
    return dups

"""Multi-GPU --fast mode, round 2: hash-range sharding with regions in the owners' key stores (csrc/shard2.cuh).

One process per GPU.  Nothing is staged, sized or acknowledged through the host any more: a rank's scatter kernel writes
every packed key row into a staging area laid out like the owners' key-store regions, the copy engines move every owner's
part straight into that owner's key store over mapped peer memory (NVLink / NVSwitch) while the SMs split the next chunk,
the owner inserts region by region and writes the duplicate flags straight back into the source's flag regions, and the
ORDER between the
ranks' streams is carried by interprocess CUDA events.  The host loop below only enqueues; its one barrier per chunk
makes sure an event has been recorded (enqueued) by its owner before a peer enqueues the wait for it.

    pack(0) pack(1) |  for c in 0 .. n-1:  insert(c)  |  apply(c)  [pack(c+2)]            finish

`ops` is any object with the methods of GpuShard2Ops and `barrier` any callable, which is how
tests/test_sharded2_cpu.py drives the same loop with gloo ranks and a CPU stand-in.
"""
from __future__ import annotations

import ctypes as C


class GpuShard2Ops:
    """One rank's engine through the C ABI (fqd_shard2_*)."""

    def __init__(self, pkg, engine, world, rank, region_rows):
        self.pkg, self.eng, self.world, self.rank = pkg, engine, world, rank
        lib = self.lib = pkg.load_library()
        vp, u64 = C.c_void_p, C.c_uint64
        lib.fqd_shard2_blob_bytes.restype = C.c_size_t
        lib.fqd_shard2_init.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.fqd_shard2_export.argtypes = [vp, vp]
        lib.fqd_shard2_import.argtypes = [vp, C.c_uint32, vp]
        lib.fqd_shard2_pack.argtypes = [vp, u64, vp, C.c_size_t, vp, C.c_size_t]
        lib.fqd_shard2_insert.argtypes = [vp, u64]
        lib.fqd_shard2_apply.argtypes = [vp, u64]
        lib.fqd_shard2_finish.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        lib.fqd_shard2_reset.argtypes = [vp]
        lib.fqd_shard2_read_flags.argtypes = [vp, vp, C.c_size_t]
        lib.fqd_shard2_timer_start.argtypes = [vp]
        lib.fqd_shard2_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
        engine._check(lib.fqd_shard2_init(engine.h, world, rank, region_rows))

    def export(self) -> bytes:
        blob = C.create_string_buffer(int(self.lib.fqd_shard2_blob_bytes()))
        self.eng._check(self.lib.fqd_shard2_export(self.eng.h, blob))
        return blob.raw

    def import_peer(self, rank: int, blob: bytes):
        self.eng._check(self.lib.fqd_shard2_import(self.eng.h, rank, blob))

    def pack(self, chunk, ptr1, n1, ptr2=None, n2=0):
        self.eng._check(self.lib.fqd_shard2_pack(self.eng.h, chunk, C.c_void_p(ptr1), n1, C.c_void_p(ptr2) if ptr2 else None, n2))

    def insert(self, chunk):
        self.eng._check(self.lib.fqd_shard2_insert(self.eng.h, chunk))

    def apply(self, chunk):
        self.eng._check(self.lib.fqd_shard2_apply(self.eng.h, chunk))

    def finish(self):
        n, d = C.c_uint64(0), C.c_uint64(0)
        self.eng._check(self.lib.fqd_shard2_finish(self.eng.h, C.byref(n), C.byref(d)))
        st = self.eng.stats()
        if st.err:
            raise self.pkg.FqdError(st.err, (self.lib.fqd_last_error(self.eng.h) or b"").decode() or f"data error at record {st.err_record} of this rank")
        return int(n.value), int(d.value)

    def reset(self):
        self.eng._check(self.lib.fqd_shard2_reset(self.eng.h))

    def timer_start(self):
        self.eng._check(self.lib.fqd_shard2_timer_start(self.eng.h))

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        self.eng._check(self.lib.fqd_shard2_timer_stop(self.eng.h, C.byref(ms)))
        return ms.value

    def read_flags(self, n):
        out = C.create_string_buffer(int(n))
        self.eng._check(self.lib.fqd_shard2_read_flags(self.eng.h, out, int(n)))
        return out.raw


def connect(ops, dist, rank, world):
    """Every rank maps every other rank's key store, hash / count / flag regions and events (CUDA IPC)."""
    blobs = [None] * world
    dist.all_gather_object(blobs, ops.export())
    for r in range(world):
        if r != rank:
            ops.import_peer(r, blobs[r])
    dist.barrier()


class HostBarrier:
    """Shared-memory barrier between the rank processes of one box (fqd_hostbar_*): a few microseconds instead of the
    0.2 - 1 ms of a gloo / NCCL barrier.  Collective constructor: every rank calls it at the same point."""

    def __init__(self, pkg, dist, rank, world):
        import os
        self.lib = pkg.load_library()
        self.lib.fqd_hostbar_open.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
        self.lib.fqd_hostbar_wait.argtypes = [C.c_void_p]
        self.lib.fqd_hostbar_close.argtypes = [C.c_void_p, C.c_char_p]
        names = [f"/fqd_bar_{os.getpid()}_{os.urandom(4).hex()}"]
        dist.broadcast_object_list(names, src=0)
        self.name, self.rank = names[0].encode(), rank
        self.h = C.c_void_p()
        assert self.lib.fqd_hostbar_open(self.name, world, C.byref(self.h)) == 0
        dist.barrier()                      # everybody has mapped it

    def __call__(self):
        self.lib.fqd_hostbar_wait(self.h)

    def close(self):
        if self.h:
            self.lib.fqd_hostbar_close(self.h, self.name if self.rank == 0 else None)
            self.h = None


def region_rows_for(chunk_records: int, world: int) -> int:
    """Rows one source sends one owner per chunk: chunk / world on average (the hash spreads keys evenly, binomial spread
    ~ sqrt of that), plus a margin that no plausible input reaches; an overflow is detected and fails the job."""
    mean = chunk_records / world
    return int(mean + 8 * mean ** 0.5 + mean * 0.03) + 4096


def run_job(ops, barrier, chunks, flags_out=None, records=None):
    """All chunks of one job.  chunks: [(ptr1, n1)] or [(ptr1, n1, ptr2, n2)], the same count on every rank (a rank whose
    slice is shorter passes empty chunks).  Returns (records of this rank, duplicates among them).
    flags_out + records (records per chunk): the per-record duplicate flags of every chunk are appended (tests)."""
    n = len(chunks)
    for c in range(min(2, n)):                        # two chunks ahead: the GPUs never wait for the host
        ops.pack(c, *chunks[c])
    barrier()
    for c in range(n):
        ops.insert(c)
        barrier()
        ops.apply(c)                                  # before the pack that reuses this chunk's parity
        if flags_out is not None:                     # per-record duplicate flags of this chunk, in input order (tests)
            flags_out.append(ops.read_flags(records[c]))
        if c + 2 < n:
            ops.pack(c + 2, *chunks[c + 2])
    return ops.finish()

"""fastq-dupaway_b200 - Python binding (ctypes) of libfqd_cuda.so, the B200 deduplication engine.

This module is test / benchmark plumbing over the C ABI declared in include/fqd.h; the drop-in product is the
C++ command line in fastq-dupaway_b200/host (same flags as the reference, src/main.cpp:40-179).  It mirrors the
reference's in-process seam (HashDupRemover / SeqDupRemover, src/hash_dup_remover.hpp:73-94,
src/seq_dup_remover.hpp:12-38) on byte buffers so that parity tests read like the reference's own tests.

There is no CPU fallback: if libfqd_cuda.so is missing or no CUDA device is usable, calls raise.
Import with importlib.import_module("fastq-dupaway_b200") (the directory name is not a Python identifier).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
import os as _os
# FQD_LIB: an experiment build of the same library (csrc/Makefile `variants`); never a different implementation
LIB_PATH = Path(_os.environ["FQD_LIB"]).resolve() if _os.environ.get("FQD_LIB") else CSRC / "libfqd_cuda.so"
HEADER = HERE.parent / "include" / "fqd.h"

FQD_ABI_VERSION = 1
FORMAT_FASTQ, FORMAT_FASTA = 0, 1
MODE_FAST, MODE_SEQ_TIGHT, MODE_SEQ_LOOSE, MODE_SEQ_HAMMING = 0, 1, 2, 3
MODE_BY_NAME = {"fast": MODE_FAST, "tight": MODE_SEQ_TIGHT, "loose": MODE_SEQ_LOOSE, "tail-hamming": MODE_SEQ_HAMMING}

STATUS = {0: "FQD_OK", 1: "FQD_ERR_INVALID", 2: "FQD_ERR_CUDA", 3: "FQD_ERR_EMPTY", 4: "FQD_ERR_BAD_START",
          5: "FQD_ERR_LEN_MISMATCH", 6: "FQD_ERR_BAD_BASE", 7: "FQD_ERR_CAPACITY", 8: "FQD_ERR_SEQ_TOO_LONG", 9: "FQD_ERR_UNSUPPORTED_BYTE", 10: "FQD_ERR_TAG_TOO_LONG"}


NONE64 = 0xFFFFFFFFFFFFFFFF


class FqdError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("mode", C.c_int32), ("format", C.c_int32),
                ("paired", C.c_int32), ("unordered", C.c_int32), ("hamming_dist", C.c_uint32),
                ("max_seq_len", C.c_uint32), ("max_records", C.c_uint64), ("max_chunk_bytes", C.c_uint64),
                ("max_chunk_records", C.c_uint64), ("max_tag_len", C.c_uint32), ("byte_keys", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("total", C.c_uint64), ("dups", C.c_uint64), ("unmatched", C.c_uint64),
                ("err", C.c_int32), ("err_char", C.c_int32), ("err_record", C.c_uint64),
                ("err_mate", C.c_int32), ("reserved", C.c_int32)]


class ChunkResult(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("consumed", C.c_uint64 * 2), ("first_record", C.c_uint64),
                ("n_survivors", C.c_uint64), ("rec_start", C.POINTER(C.c_uint32) * 2), ("dup", C.POINTER(C.c_uint8))]


class Emission(C.Structure):
    _fields_ = [("n_out", C.c_uint64), ("off", C.POINTER(C.c_uint64) * 2), ("len", C.POINTER(C.c_uint32) * 2),
                ("head", C.POINTER(C.c_uint64))]


class Profile(C.Structure):
    _fields_ = [("parse_ms", C.c_double), ("parse_launches", C.c_uint64), ("parse_bytes", C.c_uint64),
                ("parse_records", C.c_uint64), ("insert_ms", C.c_double), ("insert_launches", C.c_uint64),
                ("scatter_ms", C.c_double), ("scatter_launches", C.c_uint64)]


_lib = None


def build(verbose: bool = False) -> Path:
    """Compile libfqd_cuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", str(CSRC), "all"], capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode:
        raise RuntimeError("building libfqd_cuda.so failed")
    return LIB_PATH


def load_library():
    """Load libfqd_cuda.so; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(f"{LIB_PATH} is missing - run __graft_entry__.build() / make -C {CSRC}")
    lib = C.CDLL(str(LIB_PATH))
    vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
    lib.fqd_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.fqd_destroy.argtypes = [vp]
    lib.fqd_destroy.restype = None
    lib.fqd_last_error.argtypes = [vp]
    lib.fqd_last_error.restype = C.c_char_p
    lib.fqd_host_alloc.argtypes = [C.POINTER(vp), sz]
    lib.fqd_host_free.argtypes = [vp]
    lib.fqd_push.argtypes = [vp, vp, sz, vp, sz, C.POINTER(ChunkResult)]
    lib.fqd_push_prefetch.argtypes = [vp, vp, sz, vp, sz]
    lib.fqd_push_staged.argtypes = [vp, C.POINTER(ChunkResult)]
    lib.fqd_push_device.argtypes = [vp, vp, sz, vp, sz, C.POINTER(ChunkResult)]
    lib.fqd_push_device_async.argtypes = [vp, vp, sz, vp, sz]
    lib.fqd_sync.argtypes = [vp]
    lib.fqd_reset.argtypes = [vp]
    lib.fqd_keep_survivors.argtypes = [vp, C.c_int]
    lib.fqd_survivors.argtypes = [vp, u64, vp, u64, C.POINTER(u64), C.POINTER(vp)]
    lib.fqd_timer_start.argtypes = [vp]
    lib.fqd_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
    lib.fqd_profile_enable.argtypes = [vp, C.c_int]
    lib.fqd_profile_get.argtypes = [vp, C.POINTER(Profile)]
    lib.fqd_append.argtypes = [vp, C.c_int, vp, sz]
    lib.fqd_append_device.argtypes = [vp, C.c_int, vp, sz]
    lib.fqd_adopt_device.argtypes = [vp, C.c_int, vp, sz]
    lib.fqd_finish.argtypes = [vp]
    lib.fqd_finish_scan.argtypes = [vp]
    lib.fqd_finish_emit.argtypes = [vp]
    lib.fqd_boundary_bytes.argtypes = [vp]
    lib.fqd_boundary_bytes.restype = sz
    lib.fqd_boundary_get.argtypes = [vp, vp]
    lib.fqd_boundary_fix.argtypes = [vp, vp]
    lib.fqd_partition_sample.argtypes = [vp, C.c_uint32, C.POINTER(u64), C.POINTER(u64)]
    lib.fqd_partition_plan.argtypes = [vp, C.POINTER(u64), C.c_uint32, C.POINTER(u64), C.POINTER(u64)]
    lib.fqd_partition_gather.argtypes = [vp, C.c_int, vp]
    lib.fqd_unordered_prepare.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    lib.fqd_unordered_enter.argtypes = [vp, C.c_int, u64, C.POINTER(u64)]
    lib.fqd_unordered_join.argtypes = [vp, u64, u64, u64, u64, C.POINTER(u64)]
    lib.fqd_unordered_row_bytes.argtypes = [vp]
    lib.fqd_unordered_row_bytes.restype = sz
    lib.fqd_unordered_rows.argtypes = [vp, u64, C.c_uint32, vp, C.POINTER(u64)]
    lib.fqd_unordered_insert.argtypes = [vp, vp, u64, C.c_uint32, vp]
    lib.fqd_unordered_apply.argtypes = [vp, vp, u64, C.c_int]
    lib.fqd_emission.argtypes = [vp, C.POINTER(Emission)]
    lib.fqd_emit.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz), C.POINTER(C.c_int)]
    lib.fqd_emit_clusters.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz), C.POINTER(C.c_int)]
    lib.fqd_discard_input.argtypes = [vp, C.c_int]
    lib.fqd_emission_count.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    lib.fqd_emission_read.argtypes = [vp, C.c_int, u64, u64, vp, vp]
    lib.fqd_cluster_read.argtypes = [vp, C.c_int, u64, u64, vp, vp, vp]
    lib.fqd_device_memory.argtypes = [C.c_int, C.POINTER(sz), C.POINTER(sz)]
    lib.fqd_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.fqd_set_stream.argtypes = [vp, vp]
    lib.fqd_shard_row_bytes.argtypes = [vp]
    lib.fqd_shard_row_bytes.restype = sz
    lib.fqd_shard_pack.argtypes = [vp, vp, sz, C.c_uint32, vp, C.POINTER(u64), C.POINTER(u64)]
    lib.fqd_shard_pack_pe.argtypes = [vp, vp, sz, vp, sz, C.c_uint32, vp, C.POINTER(u64), C.POINTER(u64)]
    lib.fqd_shard_insert.argtypes = [vp, vp, u64, C.c_uint32, vp]
    lib.fqd_shard_apply.argtypes = [vp, vp, C.POINTER(u64)]
    lib.fqd_shard_read_flags.argtypes = [vp, vp, sz]
    lib.fqd_device_time_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64)]
    lib.fqd_synth_fastq.argtypes = [C.c_int, vp, u64, u64, C.c_uint32, C.c_int, u64, C.c_uint32, C.c_uint32, C.c_int]
    lib.fqd_synth_record_bytes.argtypes = [C.c_uint32]
    lib.fqd_synth_record_bytes.restype = sz
    lib.fqd_ipc_export.argtypes = [C.c_int, vp, vp]
    lib.fqd_ipc_open.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.fqd_ipc_close.argtypes = [C.c_int, vp]
    lib.fqd_peer_copy_async.argtypes = [C.c_int, vp, vp, sz, vp]
    lib.fqd_device_alloc.argtypes = [C.c_int, C.POINTER(vp), sz]
    lib.fqd_device_free.argtypes = [C.c_int, vp]
    lib.fqd_memcpy_d2h.argtypes = [C.c_int, vp, vp, sz]
    lib.fqd_memcpy_h2d.argtypes = [C.c_int, vp, vp, sz]
    _lib = lib
    return lib


def declared_symbols():
    """Function names declared in include/fqd.h (used by the CPU-side export test)."""
    import re
    txt = HEADER.read_text()
    return sorted(set(re.findall(r"^(?:int|void|size_t|const char\*)\s+(fqd_[a-z0-9_]+)\s*\(", txt, flags=re.M)))


class DeviceBuffer:
    """Raw device allocation through the C ABI (no torch needed)."""

    def __init__(self, nbytes: int, device: int = 0):
        self.lib = load_library()
        self.device = device
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        rc = self.lib.fqd_device_alloc(device, C.byref(p), max(self.nbytes, 16))
        if rc:
            raise FqdError(rc, f"cudaMalloc({nbytes})")
        self.ptr = p.value

    def upload(self, data: bytes, offset: int = 0):
        rc = self.lib.fqd_memcpy_h2d(self.device, C.c_void_p(self.ptr + offset), data, len(data))
        if rc:
            raise FqdError(rc, "h2d")

    def download(self, nbytes: int | None = None, offset: int = 0) -> bytes:
        n = self.nbytes - offset if nbytes is None else nbytes
        buf = C.create_string_buffer(n)
        rc = self.lib.fqd_memcpy_d2h(self.device, buf, C.c_void_p(self.ptr + offset), n)
        if rc:
            raise FqdError(rc, "d2h")
        return buf.raw

    def free(self):
        if self.ptr:
            self.lib.fqd_device_free(self.device, C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One fqd_handle."""

    def __init__(self, mode="fast", fmt=FORMAT_FASTQ, paired=False, unordered=False, hamming_dist=2,
                 max_seq_len=150, max_records=1 << 20, max_chunk_bytes=64 << 20, max_chunk_records=0, device=0,
                 max_tag_len=0, byte_keys=0):
        self.lib = load_library()
        cfg = Config(FQD_ABI_VERSION, device, MODE_BY_NAME[mode] if isinstance(mode, str) else mode, fmt,
                     int(paired), int(unordered), hamming_dist, max_seq_len, max_records, max_chunk_bytes,
                     max_chunk_records, max_tag_len, int(byte_keys))
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.lib.fqd_create(C.byref(cfg), C.byref(self.h))
        if rc:
            raise FqdError(rc, (self.lib.fqd_last_error(None) or b"").decode())

    def _check(self, rc):
        if rc:
            raise FqdError(rc, (self.lib.fqd_last_error(self.h) or b"").decode())

    def close(self):
        if self.h:
            self.lib.fqd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- fast ordered mode
    def push(self, r1: bytes, r2: bytes | None = None) -> ChunkResult:
        res = ChunkResult()
        self._check(self.lib.fqd_push(self.h, r1, len(r1), r2, len(r2) if r2 is not None else 0, C.byref(res)))
        return res

    def push_device(self, d1, n1, d2=None, n2=0) -> ChunkResult:
        res = ChunkResult()
        self._check(self.lib.fqd_push_device(self.h, C.c_void_p(d1), n1, C.c_void_p(d2) if d2 else None, n2, C.byref(res)))
        return res

    def push_device_async(self, d1, n1, d2=None, n2=0):
        self._check(self.lib.fqd_push_device_async(self.h, C.c_void_p(d1), n1, C.c_void_p(d2) if d2 else None, n2))

    def sync(self):
        self._check(self.lib.fqd_sync(self.h))

    def reset(self):
        self._check(self.lib.fqd_reset(self.h))

    def keep_survivors(self, on=True):
        self._check(self.lib.fqd_keep_survivors(self.h, int(on)))

    def survivors(self, fetch=True):
        """-> (numpy uint64 array of the written records' global indices | None, count, device pointer of the list)"""
        n, dp = C.c_uint64(0), C.c_void_p()
        self._check(self.lib.fqd_survivors(self.h, 0, None, 0, C.byref(n), C.byref(dp)))
        arr = None
        if fetch:
            arr = np.empty(int(n.value), dtype=np.uint64)
            if n.value:
                self._check(self.lib.fqd_survivors(self.h, 0, arr.ctypes.data_as(C.c_void_p), int(n.value), None, None))
        return arr, int(n.value), dp.value

    def timer_start(self):
        self._check(self.lib.fqd_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        self._check(self.lib.fqd_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._check(self.lib.fqd_profile_enable(self.h, int(on)))

    def profile(self) -> Profile:
        p = Profile()
        self._check(self.lib.fqd_profile_get(self.h, C.byref(p)))
        return p

    # -- whole-input modes
    def append(self, mate: int, buf: bytes):
        self._check(self.lib.fqd_append(self.h, mate, buf, len(buf)))

    def append_device(self, mate: int, dptr: int, n: int):
        self._check(self.lib.fqd_append_device(self.h, mate, C.c_void_p(dptr), n))

    def adopt_device(self, mate: int, dptr: int, n: int):
        """Zero-copy: the whole input of `mate` is the device buffer [dptr, dptr + n) (must outlive the job)."""
        self._check(self.lib.fqd_adopt_device(self.h, mate, C.c_void_p(dptr), n))

    def finish(self):
        self._check(self.lib.fqd_finish(self.h))

    # -- staged finish + repartition by key range (multi-GPU sequence mode, see sharded_seq.py)
    def finish_scan(self):
        self._check(self.lib.fqd_finish_scan(self.h))

    def finish_emit(self):
        self._check(self.lib.fqd_finish_emit(self.h))

    def boundary_get(self) -> bytes:
        n = int(self.lib.fqd_boundary_bytes(self.h))
        buf = C.create_string_buffer(n)
        self._check(self.lib.fqd_boundary_get(self.h, buf))
        return buf.raw

    def boundary_fix(self, prev: bytes):
        assert len(prev) == int(self.lib.fqd_boundary_bytes(self.h))
        self._check(self.lib.fqd_boundary_fix(self.h, prev))

    def partition_sample(self, n_samples: int):
        """-> (numpy uint64 array [n_samples, 2] of (word 0, word 1) key pairs, records in this slice)"""
        out = (C.c_uint64 * (2 * n_samples))()
        n = C.c_uint64(0)
        self._check(self.lib.fqd_partition_sample(self.h, n_samples, out, C.byref(n)))
        return np.frombuffer(out, dtype=np.uint64).reshape(n_samples, 2).copy(), int(n.value)

    def partition_plan(self, splitters, n_ranges: int):
        """splitters: uint64 array [n_ranges - 1, 2] -> (records per owner, bytes[mate][owner])"""
        mates = 2 if self.cfg.paired else 1
        sp = np.ascontiguousarray(splitters, dtype=np.uint64).reshape(-1)
        arr = (C.c_uint64 * max(1, sp.size))(*[int(x) for x in sp])
        counts = (C.c_uint64 * n_ranges)()
        nbytes = (C.c_uint64 * (mates * n_ranges))()
        self._check(self.lib.fqd_partition_plan(self.h, arr, n_ranges, counts, nbytes))
        return [int(c) for c in counts], [[int(nbytes[m * n_ranges + o]) for o in range(n_ranges)] for m in range(mates)]

    def partition_gather(self, mate: int, dptr: int):
        self._check(self.lib.fqd_partition_gather(self.h, mate, C.c_void_p(dptr)))

    # -- --unordered by stages (tag ranges across GPUs, sharded_unordered.py)
    def unordered_prepare(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.fqd_unordered_prepare(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def unordered_enter(self, side: int, i: int) -> int:
        pos = C.c_uint64(0)
        self._check(self.lib.fqd_unordered_enter(self.h, side, i, C.byref(pos)))
        return int(pos.value)

    def unordered_join(self, limit_i: int, limit_j: int, final_i: int, final_j: int):
        """-> (pairs emitted, unmatched here, emission index of the first bad pair or None, last comparison matched here)"""
        out = (C.c_uint64 * 4)()
        self._check(self.lib.fqd_unordered_join(self.h, limit_i, limit_j, final_i, final_j, out))
        bad = int(out[2])
        return int(out[0]), int(out[1]), (None if bad == NONE64 else bad), bool(out[3])

    def unordered_rows(self, limit: int, n_shards: int, dptr: int):
        counts = (C.c_uint64 * n_shards)()
        self._check(self.lib.fqd_unordered_rows(self.h, limit, n_shards, C.c_void_p(dptr), counts))
        return [int(c) for c in counts]

    def unordered_insert(self, recv_ptr: int, n_recv: int, n_shards: int, flags_ptr: int):
        self._check(self.lib.fqd_unordered_insert(self.h, C.c_void_p(recv_ptr), n_recv, n_shards, C.c_void_p(flags_ptr)))

    def unordered_apply(self, flags_ptr: int, limit: int, report_bad: bool):
        self._check(self.lib.fqd_unordered_apply(self.h, C.c_void_p(flags_ptr), limit, int(report_bad)))

    def emission(self) -> Emission:
        em = Emission()
        self._check(self.lib.fqd_emission(self.h, C.byref(em)))
        return em

    def emit_all(self, mate: int, cap: int = 1 << 22) -> bytes:
        """Output bytes of one mate, gathered on the device (fqd_emit) in emission order."""
        buf = C.create_string_buffer(cap)
        out = []
        while True:
            n, done = C.c_size_t(0), C.c_int(0)
            self._check(self.lib.fqd_emit(self.h, mate, buf, cap, C.byref(n), C.byref(done)))
            out.append(buf.raw[: n.value])
            if done.value:
                break
        return b"".join(out)

    def emit_clusters_all(self, mate: int, cap: int = 1 << 20) -> bytes:
        """Text of `<out>.clusters` for one mate (--write-clusters), gathered on the device."""
        buf = C.create_string_buffer(cap)
        out = []
        while True:
            n, done = C.c_size_t(0), C.c_int(0)
            self._check(self.lib.fqd_emit_clusters(self.h, mate, buf, cap, C.byref(n), C.byref(done)))
            out.append(buf.raw[: n.value])
            if done.value:
                break
        return b"".join(out)

    # -- whole-input modes without the raw input in device memory (inputs larger than HBM)
    def discard_input(self, on: bool = True):
        """Free every input segment as soon as it is parsed; the caller fetches the written records itself."""
        self._check(self.lib.fqd_discard_input(self.h, int(on)))

    def emission_count(self):
        """(records written, sorted positions available to cluster_read)."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.fqd_emission_count(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def emission_read(self, mate: int, first: int, count: int):
        """(offsets uint64[count], lengths uint32[count]) of the written records [first, first + count) of `mate`."""
        import numpy as np
        off = np.empty(count, dtype=np.uint64)
        ln = np.empty(count, dtype=np.uint32)
        self._check(self.lib.fqd_emission_read(self.h, mate, first, count, off.ctypes.data, ln.ctypes.data))
        return off, ln

    def cluster_read(self, mate: int, first: int, count: int):
        """Sorted positions [first, first + count): (offsets, lengths, head flags) of the records standing there."""
        import numpy as np
        off = np.empty(count, dtype=np.uint64)
        ln = np.empty(count, dtype=np.uint32)
        head = np.empty(count, dtype=np.uint8)
        self._check(self.lib.fqd_cluster_read(self.h, mate, first, count, off.ctypes.data, ln.ctypes.data, head.ctypes.data))
        return off, ln, head

    def stats(self) -> Stats:
        st = Stats()
        self._check(self.lib.fqd_stats(self.h, C.byref(st)))
        return st

    def device_time_ms(self):
        ms, n = C.c_double(0), C.c_uint64(0)
        self._check(self.lib.fqd_device_time_ms(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def _gather_spans(buf: bytes, starts: np.ndarray, keep: np.ndarray) -> bytes:
    """Concatenate records [starts[i], starts[i+1]) for which keep[i]; merges adjacent survivors."""
    mv = memoryview(buf)
    out = []
    n = len(keep)
    i = 0
    while i < n:
        if not keep[i]:
            i += 1
            continue
        j = i
        while j + 1 < n and keep[j + 1]:
            j += 1
        out.append(mv[int(starts[i]): int(starts[j + 1])])
        i = j + 1
    return b"".join(out)


def dedup_fast(b1: bytes, b2: bytes | None = None, fmt=FORMAT_FASTQ, chunk_bytes=1 << 20, max_seq_len=150,
               max_records=None, device=0):
    """--fast ordered mode on byte buffers, streamed in chunks the way the host reader does it:
    push a chunk, write the surviving records, carry the incomplete tail into the next chunk
    (src/bufferedinput.hpp:57-88).  Returns (out1, out2 | None, Stats)."""
    paired = b2 is not None
    if max_records is None:
        max_records = max(1024, (b1.count(b"\n") // 2) + 16)
    eng = Engine("fast", fmt, paired, False, 2, max_seq_len, max_records, chunk_bytes, 0, device)
    lead = b"@" if fmt == FORMAT_FASTQ else b">"
    try:
        outs = [[], []]
        bufs = [b1, b2 if paired else b""]
        pos = [0, 0]
        n_mates = 2 if paired else 1
        first = True
        adj_total = adj_dups = 0
        held = None
        st = eng.stats()
        while True:
            chunk = [bufs[m][pos[m]: pos[m] + chunk_bytes] for m in range(n_mates)]
            if any(len(c) == 0 for c in chunk):
                if first:   # empty input: the reference's first refresh() throws (src/bufferedinput.hpp:82-85)
                    st = eng.stats()
                    st.err = 3
                    return b"", (b"" if paired else None), st
                break
            res = eng.push(chunk[0], chunk[1] if paired else None)
            n = int(res.n_records)
            st = eng.stats()
            if n == 0 and st.err == 0:
                at_end = all(pos[m] + len(chunk[m]) >= len(bufs[m]) for m in range(n_mates))
                if first and at_end:
                    st.err = 3
                    return b"", (b"" if paired else None), st
                if at_end:
                    break
                raise FqdError(7, "a single record does not fit in one chunk")
            first = False
            keep = _np(res.dup, n, np.uint8) == 0
            # A non-record byte where the NEXT record should start aborts the run while the last complete
            # record is being fetched (src/bufferedinput.hpp:90-103 pre-parses it; src/fastqview.cpp:91-92
            # checks the lead byte before looking for line ends) - that last record is then never processed.
            tail_err = None
            if st.err == 0:
                for m in range(n_mates):
                    nxt = pos[m] + int(res.consumed[m])
                    if nxt < len(bufs[m]) and bufs[m][nxt: nxt + 1] != lead:
                        tail_err = (m, bufs[m][nxt])
                        break
            if tail_err is not None:
                adj_total -= 1
                adj_dups -= int(not keep[n - 1])
                keep[n - 1] = False
            starts = [_np(res.rec_start[m], n + 1, np.int64) if n else np.zeros(1, np.int64) for m in range(n_mates)]
            # The lazy pre-parse reaches across chunks: the LAST pair of a chunk is written only once the record after it
            # has parsed, which may need the next chunk (a sequence / quality length mismatch shows when the record is
            # complete).  It is held back and written in front of the next chunk's survivors - or dropped if that chunk
            # opens with a malformed record.
            stops_here = st.err != 0 or tail_err is not None
            if held is not None and (n > 0 or stops_here):
                if not (st.err in (4, 5) and st.err_record == res.first_record):
                    for m in range(n_mates):
                        outs[m].append(held[m])
                held = None
            if n > 0 and not stops_here and keep[n - 1]:
                held = [chunk[m][int(starts[m][n - 1]): int(starts[m][n])] for m in range(n_mates)]
                keep[n - 1] = False
            for m in range(n_mates):
                outs[m].append(_gather_spans(chunk[m], starts[m], keep))
                pos[m] += int(res.consumed[m]) if st.err == 0 else 0
            if tail_err is not None:
                st.err, st.err_char, st.err_record = 4, tail_err[1], st.total
                break
            if st.err == 6 and st.err_record == 0:
                # the reference writes the very first record (pair) BEFORE it keys it (src/hash_dup_remover.hpp:118-124,
                # 216-228): a base outside {A,C,G,T,N} in record 0 still leaves record 0 in the output
                for m in range(n_mates):
                    span = _np(res.rec_start[m], 2, np.int64)
                    outs[m].append(chunk[m][int(span[0]): int(span[1])])
            if st.err:
                break
            if all(pos[m] >= len(bufs[m]) for m in range(n_mates)):
                break
        if held is not None:
            for m in range(n_mates):
                outs[m].append(held[m])
        if st.err in (0, 4):
            st.total += adj_total
            st.dups += adj_dups
        return b"".join(outs[0]), (b"".join(outs[1]) if paired else None), st
    finally:
        eng.close()


def dedup_whole(mode: str, b1: bytes, b2: bytes | None = None, fmt=FORMAT_FASTQ, dist=2, unordered=False,
                max_seq_len=150, append_bytes=1 << 22, seg_bytes=1 << 24, max_records=None, device=0, max_tag_len=0,
                device_gather=True, emit_cap=1 << 16, byte_keys=0, discard=False, window=997):
    """Sequence-based modes and --fast --unordered: whole input on the device, emission order out.
    discard=True: the raw input does not stay on the device (fqd_discard_input, inputs larger than HBM); the written
    records are fetched here from the emission list read window by window (fqd_emission_read)."""
    paired = b2 is not None
    if max_records is None:
        max_records = max(1024, max(b1.count(b"\n"), b2.count(b"\n") if paired else 0) // 2 + 16)
    eng = Engine(mode, fmt, paired, unordered, dist, max_seq_len, max_records, seg_bytes, 0, device, max_tag_len, byte_keys)
    try:
        if discard:
            eng.discard_input(True)
        for m, b in enumerate([b1, b2] if paired else [b1]):
            for o in range(0, len(b), append_bytes):
                eng.append(m, b[o: o + append_bytes])
        eng.finish()
        st = eng.stats()
        if st.err and st.err != 6:
            return b"", (b"" if paired else None), st
        em = eng.emission()
        n = int(em.n_out)
        outs = []
        for m, b in enumerate([b1, b2] if paired else [b1]):
            off = _np(em.off[m], n, np.int64)
            ln = _np(em.len[m], n, np.int64)
            mv = memoryview(b)
            outs.append(b"".join(mv[int(o): int(o + l)] for o, l in zip(off, ln)))
            if discard:            # the windowed reader must hand out the same list
                if eng.emission_count()[0] != n:
                    raise AssertionError("fqd_emission_count differs from fqd_emission")
                for k in range(0, n, window):
                    c = min(window, n - k)
                    o2, l2 = eng.emission_read(m, k, c)
                    if not (np.array_equal(o2.astype(np.int64), off[k: k + c]) and np.array_equal(l2.astype(np.int64), ln[k: k + c])):
                        raise AssertionError("fqd_emission_read differs from fqd_emission")
                continue
            if device_gather:      # the device-side gather must produce the same bytes
                got = eng.emit_all(m, emit_cap)
                if got != outs[-1]:
                    raise AssertionError("fqd_emit bytes differ from the (offset, length) emission list")
        return outs[0], (outs[1] if paired else None), st
    finally:
        eng.close()

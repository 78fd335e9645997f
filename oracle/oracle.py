"""ctypes front-end of the CPU oracle (oracle/fqd_oracle.c) and of the compiled reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke().  The product package must never import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libfqd_oracle.so"
REF_BIN = HERE / "_ref" / "fastq-dupaway"
REF_STABLE_BIN = HERE / "_ref" / "fastq-dupaway-stable"

FASTQ, FASTA = 0, 1
TIGHT, LOOSE, HAMMING = 1, 2, 3
MODE_BY_NAME = {"tight": TIGHT, "loose": LOOSE, "tail-hamming": HAMMING}

ERR_NAMES = {0: "ok", 1: "empty", 2: "bad_start", 3: "len_mismatch", 4: "bad_base", 5: "nomem"}


class Stats(C.Structure):
    _fields_ = [("total", C.c_uint64), ("dups", C.c_uint64), ("unmatched", C.c_uint64),
                ("err", C.c_int32), ("err_char", C.c_int32), ("err_record", C.c_uint64)]


_lib = None


def build():
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "all"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
        _lib = C.CDLL(str(LIB_PATH))
        u64p = C.POINTER(C.c_uint64)
        _lib.fqdo_fast_se.argtypes = [C.c_char_p, C.c_int64, C.c_int, u64p, u64p, C.POINTER(Stats)]
        _lib.fqdo_fast_pe.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, u64p, u64p, C.POINTER(Stats)]
        _lib.fqdo_fast_pe_unordered.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, u64p, u64p, u64p, C.POINTER(Stats)]
        _lib.fqdo_seq.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_uint32,
                                  u64p, u64p, u64p, u64p, C.POINTER(Stats)]
        _lib.fqdo_split.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_uint64, u64p, C.POINTER(Stats)]
        _lib.fqdo_seq2hash.argtypes = [C.c_char_p, C.c_int64, u64p, C.POINTER(C.c_int32)]
        _lib.fqdo_seq2hash.restype = C.c_int64
    return _lib


def _u64(n):
    return np.zeros(max(int(n), 1), dtype=np.uint64)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _max_records(buf: bytes) -> int:
    return buf.count(b"\n") // 2 + 2


def split(buf: bytes, fmt: int, with_id: bool = False):
    """Record table (n x 7 int64: start, idlen, seqlen, f3len, quallen, tag_off, tag_len)."""
    cap = _max_records(buf)
    table = np.zeros((cap, 7), dtype=np.int64)
    cnt = C.c_uint64(0)
    st = Stats()
    lib().fqdo_split(buf, len(buf), fmt, int(with_id), table.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(cnt), C.byref(st))
    return table[: cnt.value].copy(), st


def seq2hash(seq: bytes):
    out = _u64(len(seq) // 17 + 2)
    bad = C.c_int32(0)
    n = lib().fqdo_seq2hash(seq, len(seq), _p(out), C.byref(bad))
    if n < 0:
        raise ValueError(f"unknown character {chr(bad.value)!r}")
    return out[:n].copy()


def fast_se(buf: bytes, fmt: int):
    out = _u64(_max_records(buf))
    n = C.c_uint64(0)
    st = Stats()
    lib().fqdo_fast_se(buf, len(buf), fmt, _p(out), C.byref(n), C.byref(st))
    return out[: n.value].copy(), st


def fast_pe(b1: bytes, b2: bytes, fmt: int):
    out = _u64(min(_max_records(b1), _max_records(b2)))
    n = C.c_uint64(0)
    st = Stats()
    lib().fqdo_fast_pe(b1, len(b1), b2, len(b2), fmt, _p(out), C.byref(n), C.byref(st))
    return out[: n.value].copy(), st


def fast_pe_unordered(b1: bytes, b2: bytes, fmt: int):
    cap = max(_max_records(b1), _max_records(b2))
    o1, o2 = _u64(cap), _u64(cap)
    n = C.c_uint64(0)
    st = Stats()
    lib().fqdo_fast_pe_unordered(b1, len(b1), b2, len(b2), fmt, _p(o1), _p(o2), C.byref(n), C.byref(st))
    return o1[: n.value].copy(), o2[: n.value].copy(), st


def seq_mode(b1: bytes, b2: bytes | None, fmt: int, mode: int, dist: int = 2, want_clusters: bool = False):
    cap = _max_records(b1)
    out, order, head = _u64(cap), _u64(cap), _u64(cap)
    n = C.c_uint64(0)
    st = Stats()
    lib().fqdo_seq(b1, len(b1), b2, len(b2) if b2 is not None else 0, fmt, mode, dist,
                   _p(out), C.byref(n), _p(order), _p(head), C.byref(st))
    if want_clusters:
        return out[: n.value].copy(), st, order[: st.total].copy(), head[: st.total].copy()
    return out[: n.value].copy(), st


def gather(buf: bytes, table: np.ndarray, idx) -> bytes:
    """Concatenate the raw spans of records `idx` (what the reference writes: src/fastqview.cpp:79-87)."""
    mv = memoryview(buf)
    sizes = table[:, 1] + table[:, 2] + table[:, 3] + table[:, 4]
    return b"".join(bytes(mv[int(table[i, 0]): int(table[i, 0] + sizes[i])]) for i in idx)


def run_oracle(mode: str, fmt: int, b1: bytes, b2: bytes | None = None, dist: int = 2, unordered: bool = False):
    """Full emulation -> (out1 bytes, out2 bytes | None, Stats).  mode in {"fast","tight","loose","tail-hamming"}."""
    if mode == "fast" and b2 is None:
        idx, st = fast_se(b1, fmt)
        t1, _ = split(b1, fmt)
        return gather(b1, t1, idx), None, st
    if mode == "fast" and not unordered:
        idx, st = fast_pe(b1, b2, fmt)
        t1, _ = split(b1, fmt)
        t2, _ = split(b2, fmt)
        return gather(b1, t1, idx), gather(b2, t2, idx), st
    if mode == "fast":
        i1, i2, st = fast_pe_unordered(b1, b2, fmt)
        t1, _ = split(b1, fmt)
        t2, _ = split(b2, fmt)
        return gather(b1, t1, i1), gather(b2, t2, i2), st
    idx, st = seq_mode(b1, b2, fmt, MODE_BY_NAME[mode], dist)
    t1, _ = split(b1, fmt)
    o1 = gather(b1, t1, idx)
    o2 = None
    if b2 is not None:
        t2, _ = split(b2, fmt)
        o2 = gather(b2, t2, idx)
    return o1, o2, st


def cluster_text(mode: str, fmt: int, b1: bytes, b2: bytes | None = None, dist: int = 2):
    """Text of `<out>.clusters` per mate (--write-clusters): one line per record in sorted order, the ID line of a
    written record or "--" + the ID line of a removed one (src/seq_dup_remover.hpp:60-62,75-76,89-101;
    src/file_utils.cpp:98-112)."""
    kept, st, order, _head = seq_mode(b1, b2, fmt, MODE_BY_NAME[mode], dist, want_clusters=True)
    written = set(int(i) for i in kept)
    outs = []
    for b in ([b1] if b2 is None else [b1, b2]):
        t, _ = split(b, fmt)
        mv = memoryview(b)
        parts = []
        for i in order:
            i = int(i)
            idline = bytes(mv[int(t[i, 0]): int(t[i, 0] + t[i, 1])])
            parts.append(idline if i in written else b"--" + idline)
        outs.append(b"".join(parts))
    return outs, st


# ------------------------------------------------------------------------------------------------------
# The compiled, unmodified reference (oracle/_ref) driven through its CLI.

def ref_available(stable: bool = False) -> bool:
    return (REF_STABLE_BIN if stable else REF_BIN).exists()


def run_ref(workdir, mode: str, fmt: int, b1: bytes, b2: bytes | None = None, dist: int = 2,
            unordered: bool = False, stable: bool = False, mem_mb: int | None = None, extra=(), taskset=None):
    """Run the reference binary in `workdir` -> (returncode, out1, out2, stdout, stderr)."""
    workdir = Path(workdir)
    workdir.mkdir(parents=True, exist_ok=True)
    ext = "fq" if fmt == FASTQ else "fa"
    in1, out1 = workdir / f"in_1.{ext}", workdir / f"out_1.{ext}"
    in1.write_bytes(b1)
    cmd = [str(REF_STABLE_BIN if stable else REF_BIN), "-i", str(in1), "-o", str(out1), "-v"]
    in2 = out2 = None
    if b2 is not None:
        in2, out2 = workdir / f"in_2.{ext}", workdir / f"out_2.{ext}"
        in2.write_bytes(b2)
        cmd += ["-u", str(in2), "-p", str(out2)]
    if fmt == FASTA:
        cmd += ["--format", "fasta"]
    if mode == "fast":
        cmd += ["--fast"]
        if unordered:
            cmd += ["--unordered"]
    else:
        cmd += ["--compare-seq", mode]
        if mode == "tail-hamming":
            cmd += ["--distance", str(dist)]
    if mem_mb:
        cmd += ["-m", str(mem_mb)]
    cmd += list(extra)
    if taskset is not None:
        cmd = ["taskset", "-c", str(taskset)] + cmd
    res = subprocess.run(cmd, cwd=workdir, capture_output=True)
    o1 = out1.read_bytes() if out1.exists() else b""
    o2 = out2.read_bytes() if (out2 is not None and out2.exists()) else None
    return res.returncode, o1, o2, res.stdout.decode(), res.stderr.decode()

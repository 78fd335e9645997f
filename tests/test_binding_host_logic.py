"""The streaming helper of the Python binding (fastq-dupaway_b200.dedup_fast: push a chunk, write survivors, carry the
tail) on the CPU: the binding is pointed at the test double of the C ABI (tests/fake_engine) for the duration of a test
and its output compared with the oracle - well-formed input at many chunk sizes, and one malformed record at every
position.  On the GPU the same Python code runs over the real library (tests/test_fast_gpu.py)."""
import ctypes as C
import sys
from pathlib import Path

import pytest

import synth

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "fake_engine"))
from build import build_fake  # noqa: E402


@pytest.fixture()
def fake_binding(fqd, monkeypatch):
    lib = C.CDLL(str(build_fake()))
    vp, sz = C.c_void_p, C.c_size_t
    lib.fqd_create.argtypes = [C.POINTER(fqd.Config), C.POINTER(vp)]
    lib.fqd_destroy.argtypes = [vp]
    lib.fqd_destroy.restype = None
    lib.fqd_last_error.argtypes = [vp]
    lib.fqd_last_error.restype = C.c_char_p
    lib.fqd_push.argtypes = [vp, vp, sz, vp, sz, C.POINTER(fqd.ChunkResult)]
    lib.fqd_stats.argtypes = [vp, C.POINTER(fqd.Stats)]
    monkeypatch.setattr(fqd, "_lib", lib)          # restored after the test: the product never loads the double
    return fqd


def _check(fqd, oracle, b1, b2, fmt, chunk):
    o1, o2, st = fqd.dedup_fast(b1, b2, fmt, chunk_bytes=chunk, max_seq_len=200)
    e1, e2, est = oracle.run_oracle("fast", fmt, b1, b2)
    assert st.err == {0: 0, 1: 3, 2: 4, 3: 5, 4: 6}[est.err]
    assert o1 == e1 and (b2 is None or o2 == e2)
    if est.err == 0:
        assert (st.total, st.dups) == (est.total, est.dups)


@pytest.mark.parametrize("chunk", [700, 4096, 50_000, 1 << 20])
def test_well_formed_input_at_many_chunk_sizes(fake_binding, oracle, chunk):
    seqs = synth.make_reads(3000, seed=60, read_len=90, var_len=True, n_frac=0.03, dup_frac=0.4)
    _check(fake_binding, oracle, synth.to_fastq(seqs), None, fake_binding.FORMAT_FASTQ, chunk)
    s1, s2 = synth.make_pair(2000, seed=61, read_len=70)
    _check(fake_binding, oracle, synth.to_fastq(s1, mate=1), synth.to_fastq(s2[:1800], mate=2), fake_binding.FORMAT_FASTQ, chunk)


@pytest.mark.parametrize("kind", ["start", "length", "base"])
def test_malformed_record_at_every_position(fake_binding, oracle, kind):
    from test_cli_host_logic import _damage, _records
    r = [_records(30, seed=62, mate=1), _records(30, seed=63, mate=2, read_len=80)]
    good = [b"".join(r[0]), b"".join(r[1])]
    for pos in range(30):
        bad = [b"".join(r[m][:pos] + [_damage(r[m][pos], kind)] + r[m][pos + 1:]) for m in (0, 1)]
        for chunk in (1000, 4096, 1 << 20):
            _check(fake_binding, oracle, bad[0], None, fake_binding.FORMAT_FASTQ, chunk)
            _check(fake_binding, oracle, bad[0], good[1], fake_binding.FORMAT_FASTQ, chunk)
            _check(fake_binding, oracle, good[0], bad[1], fake_binding.FORMAT_FASTQ, chunk)

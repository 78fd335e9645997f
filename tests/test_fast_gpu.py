"""GPU parity tests for --fast (ordered) mode: the CUDA path through the C ABI against the CPU oracle."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _fx(golden_dir, kind, name):
    return (golden_dir / "ref_fixtures" / kind / name).read_bytes()


def _check_se(fqd, oracle, buf, fmt, **kw):
    out, _, st = fqd.dedup_fast(buf, None, fmt, **kw)
    exp, _, est = oracle.run_oracle("fast", fmt, buf)
    assert st.err == {0: 0, 1: 3, 2: 4, 3: 5, 4: 6}[est.err]
    assert out == exp
    if est.err == 0:
        assert (st.total, st.dups) == (est.total, est.dups)
    return out, st


def _check_pe(fqd, oracle, b1, b2, fmt, **kw):
    o1, o2, st = fqd.dedup_fast(b1, b2, fmt, **kw)
    e1, e2, est = oracle.run_oracle("fast", fmt, b1, b2)
    assert st.err == {0: 0, 1: 3, 2: 4, 3: 5, 4: 6}[est.err]
    assert o1 == e1 and o2 == e2
    if est.err == 0:
        assert (st.total, st.dups) == (est.total, est.dups)
    return o1, o2, st


def test_reference_fixture_single_fast(fqd, golden_dir):
    # test/test_fast.py:7-26
    out, _, st = fqd.dedup_fast(_fx(golden_dir, "inputs", "single_fast.fa"), None, fqd.FORMAT_FASTA)
    assert out == _fx(golden_dir, "expected", "single_fast.fa")
    assert (st.total, st.dups) == (10, 4)


def test_reference_fixture_paired_fast(fqd, golden_dir):
    # test/test_fast.py:29-57
    o1, o2, st = fqd.dedup_fast(_fx(golden_dir, "inputs", "paired_fast_r1.fa"), _fx(golden_dir, "inputs", "paired_fast_r2.fa"),
                                fqd.FORMAT_FASTA)
    assert o1 == _fx(golden_dir, "expected", "paired_fast_r1.fa")
    assert o2 == _fx(golden_dir, "expected", "paired_fast_r2.fa")
    assert (st.total, st.dups) == (10, 3)


def test_reference_fixtures_as_fastq(fqd, oracle, golden_dir):
    # BASELINE config 1: the same records re-emitted as FASTQ with constant quals
    fa = _fx(golden_dir, "inputs", "single_fast.fa").split(b"\n")
    fq = b"".join(b"@" + fa[i][1:] + b"\n" + fa[i + 1] + b"\n+\n" + b"I" * len(fa[i + 1]) + b"\n" for i in range(0, len(fa) - 1, 2))
    _check_se(fqd, oracle, fq, fqd.FORMAT_FASTQ)


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
@pytest.mark.parametrize("chunk", [1 << 12, 50_000, 1 << 20])
def test_random_se(fqd, oracle, fmt, chunk):
    seqs = synth.make_reads(20000, seed=3, read_len=150, var_len=True, n_frac=0.05, dup_frac=0.3)
    f = fqd.FORMAT_FASTQ if fmt == "fastq" else fqd.FORMAT_FASTA
    buf = synth.to_fastq(seqs) if fmt == "fastq" else synth.to_fasta(seqs)
    _check_se(fqd, oracle, buf, f, chunk_bytes=chunk)


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
@pytest.mark.parametrize("chunk", [1 << 13, 1 << 20])
def test_random_pe(fqd, oracle, fmt, chunk):
    s1, s2 = synth.make_pair(15000, seed=9, read_len=150, var_len=True, n_frac=0.03)
    s2 = s2[:-11]
    f = fqd.FORMAT_FASTQ if fmt == "fastq" else fqd.FORMAT_FASTA
    mk = synth.to_fastq if fmt == "fastq" else synth.to_fasta
    _check_pe(fqd, oracle, mk(s1, mate=1), mk(s2, mate=2), f, chunk_bytes=chunk)


def test_fixed_length_150bp(fqd, oracle):
    seqs = synth.make_reads(30000, seed=4, read_len=150, dup_frac=0.3)
    _check_se(fqd, oracle, synth.to_fastq(seqs), fqd.FORMAT_FASTQ, chunk_bytes=1 << 21)


def test_heavy_duplication_and_races(fqd, oracle):
    # only 7 distinct keys among 50k records: every insert races on the same buckets
    rng = np.random.default_rng(5)
    base = synth.make_reads(7, seed=6, read_len=100, dup_frac=0.0)
    seqs = [base[int(k)] for k in rng.integers(0, 7, size=50000)]
    out, st = _check_se(fqd, oracle, synth.to_fastq(seqs), fqd.FORMAT_FASTQ, chunk_bytes=1 << 22)
    assert st.total - st.dups == 7


def test_tiny_and_ragged_records(fqd, oracle):
    # empty sequences, 1-base reads, lengths around the 20-base word boundary and beyond the 1 KiB halo
    lens = [0, 1, 2, 19, 20, 21, 39, 40, 41, 59, 60, 61, 150, 0, 1, 20, 300, 1200, 3000, 20, 0]
    rng = np.random.default_rng(8)
    seqs = [bytes(rng.choice(list(b"ACGTN"), size=l).astype(np.uint8)) for l in lens]
    seqs = seqs + seqs[::-1] + seqs
    _check_se(fqd, oracle, synth.to_fastq(seqs), fqd.FORMAT_FASTQ, max_seq_len=3000, chunk_bytes=1 << 16)
    _check_se(fqd, oracle, synth.to_fasta(seqs), fqd.FORMAT_FASTA, max_seq_len=3000, chunk_bytes=1 << 16)
    # thousands of minimal FASTA records in one tile (more than one pack round per tile)
    tiny = [b"", b"A", b"C", b""] * 6000
    _check_se(fqd, oracle, synth.to_fasta(tiny, ids=[b">"] * len(tiny)), fqd.FORMAT_FASTA, chunk_bytes=1 << 16)


def test_same_prefix_different_length(fqd, oracle):
    seqs = [b"ACGT" * 10, b"ACGT" * 10 + b"A", b"ACGT" * 5, b"ACGT" * 10, b"ACGT" * 5 + b"N", b"ACGT" * 5]
    out, st = _check_se(fqd, oracle, synth.to_fastq(seqs), fqd.FORMAT_FASTQ)
    assert st.dups == 2


def test_error_paths(fqd, oracle):
    cases = [b"", b"@a\nACGT\n+\nFFF\n",
             b"@a\nACGT\n+\nFFFF\n@b\nACXT\n+\nFFFF\n@c\nAAAA\n+\nFFFF\n",
             b"@a\nACGT\n+\nFFFF\n@b\nAAAA\n+\nFFFF\nxc\nAAAA\n+\nFFFF\n",
             b"@a\nACGT\n+\nFFFF\n@b\nAAAA\n+\nFFFF",
             b"@a\nACGT\n+\nFFFF\n\n",
             b"@a\nACGT\n+\nFFFF\n@b\nacgt\n+\nFFFF\n"]
    for buf in cases:
        _check_se(fqd, oracle, buf, fqd.FORMAT_FASTQ)
    _, st = _check_se(fqd, oracle, cases[2], fqd.FORMAT_FASTQ)
    assert chr(st.err_char) == "X" and st.err_record == 1


def test_bad_base_in_the_very_first_record(fqd, oracle):
    """The reference writes record 0 (pair 0) before it keys it, so a base outside {A,C,G,T,N} there stops the run with
    that record already in the output (pinned against the reference binary in tests/test_oracle.py)."""
    se = b"@a\nACXT\n+\nFFFF\n@b\nAAAA\n+\nFFFF\n"
    out, st = _check_se(fqd, oracle, se, fqd.FORMAT_FASTQ)
    assert out == b"@a\nACXT\n+\nFFFF\n" and st.err == 6 and st.err_record == 0
    good = b"@a\nACGT\n+\nFFFF\n@b\nAAAA\n+\nFFFF\n"
    o1, o2, st = _check_pe(fqd, oracle, good, se, fqd.FORMAT_FASTQ)
    assert o1 == b"@a\nACGT\n+\nFFFF\n" and o2 == b"@a\nACXT\n+\nFFFF\n"
    _check_pe(fqd, oracle, se, good, fqd.FORMAT_FASTQ)


def test_error_in_later_chunk(fqd, oracle):
    seqs = synth.make_reads(5000, seed=10, read_len=80)
    seqs[3777] = seqs[3777][:30] + b"Z" + seqs[3777][31:]
    _check_se(fqd, oracle, synth.to_fastq(seqs), fqd.FORMAT_FASTQ, chunk_bytes=1 << 15)


def test_device_generator_roundtrip(fqd, oracle):
    """The counter-based device generator feeds fqd_push_device; its bytes, copied back, go through the oracle."""
    lib = fqd.load_library()
    n, L = 40000, 150
    rb = lib.fqd_synth_record_bytes(L)
    assert rb == 322
    dbuf = fqd.DeviceBuffer(n * rb + 4096)
    assert lib.fqd_synth_fastq(0, dbuf.ptr, 0, n, L, 1, 1, 300, 1, 0) == 0
    raw = dbuf.download(n * rb)
    exp_idx, est = oracle.fast_se(raw, oracle.FASTQ)
    assert est.err == 0 and est.total == n
    assert 0.2 < est.dups / n < 0.4
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, max_seq_len=L, max_records=n + 16, max_chunk_bytes=n * rb + 4096)
    res = eng.push_device(dbuf.ptr, n * rb)
    assert res.n_records == n and res.consumed[0] == n * rb
    dup = np.ctypeslib.as_array(res.dup, shape=(n,)).copy()
    assert np.array_equal(np.flatnonzero(dup == 0).astype(np.uint64), exp_idx)
    st = eng.stats()
    assert (st.total, st.dups) == (est.total, est.dups)
    eng.close()
    dbuf.free()


def test_large_synthetic_properties(fqd):
    """Size-independent properties at a size the oracle would not finish quickly: survivors + dups == total,
    a second pass over the SAME data finds everything duplicate (idempotence), async and sync pushes agree."""
    lib = fqd.load_library()
    n, L = 2_000_000, 150
    rb = lib.fqd_synth_record_bytes(L)
    dbuf = fqd.DeviceBuffer(n * rb + 4096)
    assert lib.fqd_synth_fastq(0, dbuf.ptr, 0, n, L, 1, 42, 300, 1, 0) == 0
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, max_seq_len=L, max_records=2 * n + 16, max_chunk_bytes=n * rb + 4096)
    half = (n // 2) * rb
    eng.push_device_async(dbuf.ptr, half)
    eng.push_device_async(dbuf.ptr + half, n * rb - half)
    eng.sync()
    st = eng.stats()
    assert st.err == 0 and st.total == n
    assert 0.25 < st.dups / n < 0.35
    first_dups = st.dups
    res = eng.push_device(dbuf.ptr, n * rb)      # same records again: all duplicates
    assert res.n_records == n and res.n_survivors == 0
    eng.close()
    eng2 = fqd.Engine("fast", fqd.FORMAT_FASTQ, max_seq_len=L, max_records=n + 16, max_chunk_bytes=n * rb + 4096)
    res2 = eng2.push_device(dbuf.ptr, n * rb)
    assert res2.n_records == n and n - res2.n_survivors == first_dups
    eng2.close()
    dbuf.free()


def test_device_generator_matches_cpu_twin(fqd):
    import bench_synth
    lib = fqd.load_library()
    n, L = 5000, 150
    rb = lib.fqd_synth_record_bytes(L)
    dbuf = fqd.DeviceBuffer(n * rb + 4096)
    for first in (0, 123457):
        assert lib.fqd_synth_fastq(0, dbuf.ptr, first, n, L, 1, 1, 300, 20, 0) == 0
        assert dbuf.download(n * rb) == bench_synth.synth_fastq_cpu(first, n, L, 1, 1, 300, 20)
    dbuf.free()


def test_prefetched_pushes_match_plain_pushes(fqd, oracle):
    """fqd_push_prefetch / fqd_push_staged (copy of chunk c+1 under the processing of chunk c): same records, same flags
    as the oracle; paired input; the chunks are cut at record boundaries."""
    import ctypes as C
    s1, s2 = synth.make_pair(9000, seed=77, read_len=90, var_len=True, n_frac=0.03, dup_frac=0.4)
    r1 = [synth.to_fastq([s], ids=[b"@q.%d 1" % i]) for i, s in enumerate(s1)]
    r2 = [synth.to_fastq([s], ids=[b"@q.%d 2" % i]) for i, s in enumerate(s2)]
    per = 1500
    chunks = [(b"".join(r1[i: i + per]), b"".join(r2[i: i + per])) for i in range(0, len(r1), per)]
    maxb = max(max(len(a), len(b)) for a, b in chunks) + 4096
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, True, False, 2, 90, 20000, maxb, 4000, 0)
    lib = eng.lib
    try:
        dup = []
        assert lib.fqd_push_prefetch(eng.h, chunks[0][0], len(chunks[0][0]), chunks[0][1], len(chunks[0][1])) == 0
        for c in range(len(chunks)):
            if c + 1 < len(chunks):
                a, b = chunks[c + 1]
                assert lib.fqd_push_prefetch(eng.h, a, len(a), b, len(b)) == 0
            res = fqd.ChunkResult()
            assert lib.fqd_push_staged(eng.h, C.byref(res)) == 0
            assert res.n_records == per
            dup.append(np.ctypeslib.as_array(res.dup, shape=(per,)).copy())
        keep_idx, est = oracle.fast_pe(b"".join(r1), b"".join(r2), oracle.FASTQ)
        got = np.concatenate(dup)
        assert np.array_equal(np.flatnonzero(got == 0).astype(np.uint64), keep_idx)
        st = eng.stats()
        assert (st.total, st.dups) == (est.total, est.dups)
    finally:
        eng.close()


@pytest.mark.parametrize("paired", [False, True])
def test_key_store_grows_in_place(fqd, oracle, paired):
    """max_records far below the input (what the host passes for a pipe, whose size it cannot know): the engine grows the
    key store and rehashes the table between chunks instead of failing - several times in a row (round-1: restart)."""
    if paired:
        s1, s2 = synth.make_pair(30000, seed=41, read_len=100, dup_frac=0.3)
        _check_pe(fqd, oracle, synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2), fqd.FORMAT_FASTQ, chunk_bytes=1 << 17, max_records=700)
    else:
        seqs = synth.make_reads(60000, seed=42, read_len=150, var_len=True, n_frac=0.02, dup_frac=0.4)
        _check_se(fqd, oracle, synth.to_fastq(seqs), fqd.FORMAT_FASTQ, chunk_bytes=1 << 18, max_records=1000)


def test_async_pushes_keep_the_first_error(fqd):
    """fqd_push_device_async does not read anything back per chunk; a data error in ANY chunk must still be reported by
    fqd_sync (round 1 looked at the last chunk only)."""
    seqs = synth.make_reads(9000, seed=43, read_len=100)
    seqs[1234] = seqs[1234][:50] + b"Z" + seqs[1234][51:]
    buf = synth.to_fastq(seqs)
    cut = [0]
    for _ in range(2):                          # three chunks, cut at record boundaries
        cut.append(buf.index(b"\n@SYN.", cut[-1] + len(buf) // 4) + 1)
    cut.append(len(buf))
    dbuf = fqd.DeviceBuffer(len(buf) + 4096)
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, max_seq_len=100, max_records=10000, max_chunk_bytes=len(buf) + 4096)
    off = 0
    parts = []
    for a, b in zip(cut[:-1], cut[1:]):
        o = (off + 15) & ~15
        dbuf.upload(buf[a:b], o)
        parts.append((o, b - a))
        off = o + (b - a)
    for o, n in parts:
        eng.push_device_async(dbuf.ptr + o, n)
    eng.sync()
    st = eng.stats()
    assert st.err == 6 and st.err_record == 1234 and st.err_char == ord("Z")
    eng.close()
    dbuf.free()


@pytest.mark.parametrize("grow", [False, True])
def test_survivor_list_on_the_device(fqd, oracle, grow):
    """fqd_keep_survivors: the ascending list of written records over a whole multi-chunk job equals the oracle's output
    index list - with the chunks pushed asynchronously (nothing read back per chunk), and across an in-place growth."""
    lib = fqd.load_library()
    n, L, chunk = 50000, 150, 12000
    rb = lib.fqd_synth_record_bytes(L)
    dbuf = fqd.DeviceBuffer(n * rb + 4096)
    assert lib.fqd_synth_fastq(0, dbuf.ptr, 0, n, L, 1, 7, 300, 1, 0) == 0
    exp_idx, est = oracle.fast_se(dbuf.download(n * rb), oracle.FASTQ)
    eng = fqd.Engine("fast", fqd.FORMAT_FASTQ, max_seq_len=L, max_records=(5000 if grow else n + 16), max_chunk_bytes=chunk * rb + 4096)
    eng.keep_survivors(True)
    for first in range(0, n, chunk):
        cnt = min(chunk, n - first)
        if grow:
            eng.push_device(dbuf.ptr + first * rb, cnt * rb)
        else:
            eng.push_device_async(dbuf.ptr + first * rb, cnt * rb)
    eng.sync()
    got, cnt, dptr = eng.survivors()
    st = eng.stats()
    assert st.err == 0 and (st.total, st.dups) == (est.total, est.dups)
    assert cnt == est.total - est.dups and dptr
    assert np.array_equal(got, exp_idx)
    # a second job on the same handle starts a new list
    eng.reset()
    eng.push_device_async(dbuf.ptr, chunk * rb)
    eng.sync()
    got2, _, _ = eng.survivors()
    assert np.array_equal(got2, exp_idx[exp_idx < chunk])
    eng.close()
    dbuf.free()

"""GPU parity tests for --fast --unordered (device tag sort + merge-join + pair set) against the CPU oracle, which is
pinned to the reference's fixtures and binary at sizes where the reference is valid."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu

ERRMAP = {0: 0, 1: 3, 2: 4, 3: 5, 4: 6}


def _fx(golden_dir, kind, name):
    return (golden_dir / "ref_fixtures" / kind / name).read_bytes()


def _check(fqd, oracle, b1, b2, fmt, **kw):
    o1, o2, st = fqd.dedup_whole("fast", b1, b2, fmt, unordered=True, **kw)
    e1, e2, est = oracle.run_oracle("fast", fmt, b1, b2, unordered=True)
    assert st.err == ERRMAP[est.err]
    if est.err == 0:
        assert o1 == e1 and o2 == e2
        assert (st.total, st.dups, st.unmatched) == (est.total, est.dups, est.unmatched)
    return st


@pytest.mark.parametrize("name", ["shuffled", "skewed", "deletion", "interleaved", "not_overlapped"])
def test_reference_fixtures(fqd, golden_dir, name):
    # test/test_unordered.py:7-48
    o1, o2, _ = fqd.dedup_whole("fast", _fx(golden_dir, "inputs", f"unordered_{name}_r1.fa"), _fx(golden_dir, "inputs", f"unordered_{name}_r2.fa"),
                                fqd.FORMAT_FASTA, unordered=True)
    assert o1 == _fx(golden_dir, "expected", f"unordered_{name}_r1.fa")
    assert o2 == _fx(golden_dir, "expected", f"unordered_{name}_r2.fa")


def _make(n, seed, drop1=0.1, drop2=0.1, id_fmt="@RUN.{i:06d}", desc=True, shuffle=True):
    rng = np.random.default_rng(seed)
    s1, s2 = synth.make_pair(n, seed=seed, read_len=40)
    ids = [id_fmt.format(i=i).encode() for i in range(n)]
    a = [(ids[i] + (b" 1" if desc else b""), s1[i]) for i in range(n) if rng.random() > drop1]
    b = [(ids[i] + (b" 2" if desc else b""), s2[i]) for i in range(n) if rng.random() > drop2]
    if shuffle:
        rng.shuffle(b)
        rng.shuffle(a)
    return (synth.to_fastq([x[1] for x in a], ids=[x[0] for x in a]), synth.to_fastq([x[1] for x in b], ids=[x[0] for x in b]))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_shuffled_with_deletions(fqd, oracle, seed):
    b1, b2 = _make(3000, seed)
    _check(fqd, oracle, b1, b2, fqd.FORMAT_FASTQ, max_seq_len=40, seg_bytes=1 << 16)


def test_tags_without_dot_and_without_description(fqd, oracle):
    # tag = everything after '@' up to the end of line INCLUDING the newline; lexicographic, not numeric order
    b1, b2 = _make(500, 5, id_fmt="@read{i}", desc=False)
    _check(fqd, oracle, b1, b2, fqd.FORMAT_FASTQ, max_seq_len=40)


def test_dot_inside_description(fqd, oracle):
    # "@r7 len=1.5" -> tag "5\n" for every record: all tags collide, pairs are matched by rank
    seqs1 = synth.make_reads(50, seed=6, read_len=30, dup_frac=0.2)
    seqs2 = synth.make_reads(40, seed=7, read_len=30, dup_frac=0.2)
    b1 = synth.to_fastq(seqs1, ids=[f"@r{i} len=1.5".encode() for i in range(50)])
    b2 = synth.to_fastq(seqs2, ids=[f"@r{i} len=1.5".encode() for i in range(40)])
    _check(fqd, oracle, b1, b2, fqd.FORMAT_FASTQ, max_seq_len=30)


def test_duplicate_tags(fqd, oracle):
    rng = np.random.default_rng(8)
    s1 = synth.make_reads(400, seed=9, read_len=30, dup_frac=0.2)
    s2 = synth.make_reads(380, seed=10, read_len=30, dup_frac=0.2)
    ids1 = [f"@T.{int(rng.integers(0, 120)):04d} 1".encode() for _ in range(400)]
    ids2 = [f"@T.{int(rng.integers(0, 120)):04d} 2".encode() for _ in range(380)]
    _check(fqd, oracle, synth.to_fastq(s1, ids=ids1), synth.to_fastq(s2, ids=ids2), fqd.FORMAT_FASTQ, max_seq_len=30)


def test_end_of_stream_rule(fqd, oracle):
    # R1 ids {1,2,9}, R2 ids {1,2,3,9} -> pair 9 is NOT written
    def fq(ids):
        return b"".join(b"@X.%d\nACGT\n+\nFFFF\n" % i for i in ids)
    st = _check(fqd, oracle, fq([1, 2, 9]), fq([1, 2, 3, 9]), fqd.FORMAT_FASTQ)
    assert st.total == 2
    for a, b in [([1], [1]), ([1], [2]), ([1, 2], [2]), ([2], [1, 2]), ([1, 2, 3], [3]), ([5], [1, 2, 3, 4, 5]), ([1, 2, 3, 4, 5], [5]),
                 ([1, 3, 5, 7], [2, 4, 6, 8]), ([1, 2, 3], [1, 2, 3])]:
        _check(fqd, oracle, fq(a), fq(b), fqd.FORMAT_FASTQ)


def test_exhaustive_small_cases(fqd, oracle):
    # every pair of tag multisets over a tiny alphabet: the stop rule and rank pairing in all corner cases
    import itertools
    def fa(ids):
        return b"".join(b">X.%d\nAC%s\n" % (i, b"GT"[k % 2:k % 2 + 1]) for k, i in enumerate(ids))
    cases = [c for r in (1, 2, 3) for c in itertools.combinations_with_replacement([1, 2, 3], r)]
    for a in cases:
        for b in cases:
            _check(fqd, oracle, fa(a), fa(b), fqd.FORMAT_FASTA)


def test_bad_base_in_matched_pair(fqd, oracle):
    def fq(recs):
        return b"".join(b"@X.%d\n%s\n+\n%s\n" % (i, s, b"F" * len(s)) for i, s in recs)
    b1 = fq([(1, b"ACGT"), (2, b"ACXT"), (3, b"AAAA"), (4, b"CCCC")])
    b2 = fq([(1, b"ACGT"), (2, b"ACGT"), (3, b"AAAA"), (4, b"CCCC")])
    st = _check(fqd, oracle, b1, b2, fqd.FORMAT_FASTQ)
    assert st.err == 6 and chr(st.err_char) == "X"
    # the same byte in a record that is never matched is harmless
    b2u = fq([(1, b"ACGT"), (3, b"AAAA"), (4, b"CCCC"), (5, b"GGGG")])
    st = _check(fqd, oracle, b1, b2u, fqd.FORMAT_FASTQ)
    assert st.err == 0


def test_long_tags_report_their_limit(fqd):
    b = b"@" + b"Q" * 50 + b"\nACGT\n+\nFFFF\n"
    _, _, st = fqd.dedup_whole("fast", b, b, fqd.FORMAT_FASTQ, unordered=True)
    assert st.err == 10
    o1, o2, st = fqd.dedup_whole("fast", b, b, fqd.FORMAT_FASTQ, unordered=True, max_tag_len=64)
    assert st.err == 0 and o1 == b and o2 == b


@pytest.mark.parametrize("seed", [5, 6])
def test_discarded_input(fqd, oracle, seed):
    """--fast --unordered with nothing but tags, key rows and record tables resident (fqd_discard_input)."""
    b1, b2 = _make(3000, seed)
    _check(fqd, oracle, b1, b2, fqd.FORMAT_FASTQ, max_seq_len=40, seg_bytes=1 << 16, append_bytes=45_000, discard=True, window=211)

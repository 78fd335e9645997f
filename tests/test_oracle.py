"""CPU tests: pin the oracle (oracle/fqd_oracle.c) against the reference's own golden fixtures and against the
unmodified reference compiled into oracle/_ref (when present)."""
import shutil

import numpy as np
import pytest

import synth

FIX_FAST_SE = ["single_fast"]


def _fx(golden_dir, kind, name):
    return (golden_dir / "ref_fixtures" / kind / name).read_bytes()


def test_seq2hash_known_answers(oracle):
    # base-5 Horner, 17 bases per word (src/seq_utils.cpp:23-49)
    assert list(oracle.seq2hash(b"")) == []
    assert list(oracle.seq2hash(b"A")) == [0]
    assert list(oracle.seq2hash(b"ACGTN")) == [((((0 * 5 + 1) * 5 + 2) * 5 + 3) * 5 + 4)]
    w = oracle.seq2hash(b"T" * 17 + b"C")
    assert list(w) == [sum(3 * 5 ** k for k in range(17)), 1]
    with pytest.raises(ValueError):
        oracle.seq2hash(b"ACGU")


def test_fixture_fast_se(oracle, golden_dir):
    out, _, st = oracle.run_oracle("fast", oracle.FASTA, _fx(golden_dir, "inputs", "single_fast.fa"))
    assert out == _fx(golden_dir, "expected", "single_fast.fa")
    assert (st.total, st.dups) == (10, 4)


def test_fixture_fast_pe(oracle, golden_dir):
    o1, o2, st = oracle.run_oracle("fast", oracle.FASTA, _fx(golden_dir, "inputs", "paired_fast_r1.fa"),
                                   _fx(golden_dir, "inputs", "paired_fast_r2.fa"))
    assert o1 == _fx(golden_dir, "expected", "paired_fast_r1.fa")
    assert o2 == _fx(golden_dir, "expected", "paired_fast_r2.fa")
    assert (st.total, st.dups) == (10, 3)


@pytest.mark.parametrize("name,mode,dist", [("single_tight.fa", "tight", 2), ("single_loose.fa", "loose", 2),
                                            ("single_hamming.fa", "tail-hamming", 1)])
def test_fixture_seq_se(oracle, golden_dir, name, mode, dist):
    out, _, _ = oracle.run_oracle(mode, oracle.FASTA, _fx(golden_dir, "inputs", name), dist=dist)
    assert out == _fx(golden_dir, "expected", name)


def test_fixture_seq_negative_control(oracle, golden_dir):
    # test/test_seq.py:78-97: tight output differs from the hamming golden
    out, _, _ = oracle.run_oracle("tight", oracle.FASTA, _fx(golden_dir, "inputs", "single_hamming.fa"))
    assert out != _fx(golden_dir, "expected", "single_hamming.fa")


def test_fixture_seq_pe(oracle, golden_dir):
    o1, o2, _ = oracle.run_oracle("tight", oracle.FASTA, _fx(golden_dir, "inputs", "paired_tight_r1.fa"),
                                  _fx(golden_dir, "inputs", "paired_tight_r2.fa"))
    assert o1 == _fx(golden_dir, "expected", "paired_tight_r1.fa")
    assert o2 == _fx(golden_dir, "expected", "paired_tight_r2.fa")


@pytest.mark.parametrize("name", ["shuffled", "skewed", "deletion", "interleaved", "not_overlapped"])
def test_fixture_unordered(oracle, golden_dir, name):
    o1, o2, _ = oracle.run_oracle("fast", oracle.FASTA, _fx(golden_dir, "inputs", f"unordered_{name}_r1.fa"),
                                  _fx(golden_dir, "inputs", f"unordered_{name}_r2.fa"), unordered=True)
    assert o1 == _fx(golden_dir, "expected", f"unordered_{name}_r1.fa")
    assert o2 == _fx(golden_dir, "expected", f"unordered_{name}_r2.fa")


def test_oracle_error_paths(oracle):
    _, _, st = oracle.run_oracle("fast", oracle.FASTQ, b"")
    assert st.err == 1
    _, _, st = oracle.run_oracle("fast", oracle.FASTQ, b"@a\nACGT\n+\nFFF\n")
    assert st.err == 3
    out, _, st = oracle.run_oracle("fast", oracle.FASTQ, b"@a\nACGT\n+\nFFFF\n@b\nACXT\n+\nFFFF\n@c\nAAAA\n+\nFFFF\n")
    assert st.err == 4 and chr(st.err_char) == "X" and out == b"@a\nACGT\n+\nFFFF\n"
    # malformed record 2: the fetch of record 1 pre-parses it, so record 1 is never written
    out, _, st = oracle.run_oracle("fast", oracle.FASTQ, b"@a\nACGT\n+\nFFFF\n@b\nAAAA\n+\nFFFF\nxc\nAAAA\n+\nFFFF\n")
    assert st.err == 2 and out == b"@a\nACGT\n+\nFFFF\n"
    # trailing record without newline is dropped silently
    out, _, st = oracle.run_oracle("fast", oracle.FASTQ, b"@a\nACGT\n+\nFFFF\n@b\nAAAA\n+\nFFFF")
    assert st.err == 0 and out == b"@a\nACGT\n+\nFFFF\n"


# ---------------------------------------------------------------------------------------------------------
# against the compiled, unmodified reference

def _need_ref(oracle, stable=False):
    if not oracle.ref_available(stable):
        pytest.skip("oracle/_ref not built (reference sources absent)")


def _parse_counts(stdout):
    nums = [int(t) for t in stdout.replace("\n", " ").split() if t.isdigit()]
    return nums


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
@pytest.mark.parametrize("seed", [1, 2])
def test_ref_fast_se(oracle, tmp_path, fmt, seed):
    _need_ref(oracle)
    seqs = synth.make_reads(3000, seed=seed, read_len=60, var_len=True, n_frac=0.05)
    f = oracle.FASTQ if fmt == "fastq" else oracle.FASTA
    buf = synth.to_fastq(seqs) if fmt == "fastq" else synth.to_fasta(seqs)
    rc, r1, _, so, _ = oracle.run_ref(tmp_path, "fast", f, buf)
    out, _, st = oracle.run_oracle("fast", f, buf)
    assert rc == 0 and out == r1
    assert _parse_counts(so)[:2] == [st.total, st.dups]


@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
def test_ref_fast_pe(oracle, tmp_path, fmt):
    _need_ref(oracle)
    s1, s2 = synth.make_pair(2500, seed=5, read_len=40, var_len=True, n_frac=0.03)
    s2 = s2[:-7]       # files of different length: stops at the shorter one
    f = oracle.FASTQ if fmt == "fastq" else oracle.FASTA
    mk = synth.to_fastq if fmt == "fastq" else synth.to_fasta
    b1, b2 = mk(s1, mate=1), mk(s2, mate=2)
    rc, r1, r2, so, _ = oracle.run_ref(tmp_path, "fast", f, b1, b2)
    o1, o2, st = oracle.run_oracle("fast", f, b1, b2)
    assert rc == 0 and o1 == r1 and o2 == r2
    assert _parse_counts(so)[:2] == [st.total, st.dups]


@pytest.mark.parametrize("mode,dist", [("tight", 2), ("loose", 2), ("tail-hamming", 0), ("tail-hamming", 2), ("tail-hamming", 3)])
@pytest.mark.parametrize("paired", [False, True])
def test_ref_stable_seq_modes(oracle, tmp_path, mode, dist, paired):
    """Full bytes against the stable-sort build of the reference, with prefixes, Ns and tail substitutions."""
    _need_ref(oracle, stable=True)
    kw = dict(read_len=30, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    if paired:
        s1, s2 = synth.make_pair(1500, seed=11, **kw)
        b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    else:
        b1, b2 = synth.to_fastq(synth.make_reads(2000, seed=12, **kw)), None
    rc, r1, r2, so, se = oracle.run_ref(tmp_path, mode, oracle.FASTQ, b1, b2, dist=dist, stable=True, mem_mb=10240)
    o1, o2, st = oracle.run_oracle(mode, oracle.FASTQ, b1, b2, dist=dist)
    assert rc == 0, se
    assert o1 == r1 and o2 == r2
    assert _parse_counts(so)[:2] == [st.total, st.dups]


@pytest.mark.parametrize("mode,dist", [("tight", 2), ("loose", 2), ("tail-hamming", 2)])
@pytest.mark.parametrize("paired", [False, True])
def test_ref_stable_arbitrary_bytes(oracle, tmp_path, mode, dist, paired):
    """Sequence-based modes accept and order ANY byte (src/fastqview.cpp:56-67): lower case, IUPAC codes,
    gaps, a tab (sorts below the line feed).  Full bytes against the stable-sort build of the reference."""
    _need_ref(oracle, stable=True)
    kw = dict(read_len=30, var_len=True, min_len=1, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5, alphabet=b"ACGTNacgtnRYKM*-.\t")
    if paired:
        s1, s2 = synth.make_pair(1200, seed=31, **kw)
        b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    else:
        b1, b2 = synth.to_fastq(synth.make_reads(1500, seed=32, **kw)), None
    rc, r1, r2, so, se = oracle.run_ref(tmp_path, mode, oracle.FASTQ, b1, b2, dist=dist, stable=True, mem_mb=10240)
    o1, o2, st = oracle.run_oracle(mode, oracle.FASTQ, b1, b2, dist=dist)
    assert rc == 0, se
    assert o1 == r1 and o2 == r2
    assert _parse_counts(so)[:2] == [st.total, st.dups]


@pytest.mark.parametrize("mode,dist", [("tight", 2), ("loose", 2), ("tail-hamming", 2)])
@pytest.mark.parametrize("paired", [False, True])
def test_ref_stable_cluster_files(oracle, tmp_path, mode, dist, paired):
    """--write-clusters: the oracle's cluster text against the files the stable-sort build of the reference writes."""
    _need_ref(oracle, stable=True)
    kw = dict(read_len=30, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    if paired:
        s1, s2 = synth.make_pair(800, seed=21, **kw)
        b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    else:
        b1, b2 = synth.to_fastq(synth.make_reads(1000, seed=22, **kw)), None
    rc, r1, r2, so, se = oracle.run_ref(tmp_path, mode, oracle.FASTQ, b1, b2, dist=dist, stable=True, mem_mb=10240,
                                        extra=["--write-clusters"])
    assert rc == 0, se
    texts, st = oracle.cluster_text(mode, oracle.FASTQ, b1, b2, dist=dist)
    assert (tmp_path / "out_1.fq.clusters").read_bytes() == texts[0]
    if paired:
        assert (tmp_path / "out_2.fq.clusters").read_bytes() == texts[1]


def _seq_column(buf, fmt_fastq=True):
    lines = buf.split(b"\n")
    step = 4 if fmt_fastq else 2
    return lines[1::step]


@pytest.mark.parametrize("mode", ["tight", "loose", "tail-hamming"])
def test_ref_unstable_sequence_column(oracle, tmp_path, mode):
    """Against the plain (introsort) reference only the sequence column is defined."""
    _need_ref(oracle)
    seqs = synth.make_reads(3000, seed=21, read_len=25, var_len=True, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    b1 = synth.to_fastq(seqs)
    rc, r1, _, _, _ = oracle.run_ref(tmp_path, mode, oracle.FASTQ, b1, None, dist=2)
    o1, _, _ = oracle.run_oracle(mode, oracle.FASTQ, b1, None, dist=2)
    assert rc == 0
    if mode == "tail-hamming":
        # the greedy head-based scan depends on which member leads a tie group; counts still have to agree
        assert len(_seq_column(o1)) == len(_seq_column(r1))
    else:
        assert _seq_column(o1) == _seq_column(r1)


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_ref_unordered(oracle, tmp_path, seed):
    _need_ref(oracle)
    rng = np.random.default_rng(seed)
    n = 800
    s1, s2 = synth.make_pair(n, seed=seed, read_len=30)
    ids1 = [f"@RUN.{i:06d} 1".encode() for i in range(n)]
    ids2 = [f"@RUN.{i:06d} 2".encode() for i in range(n)]
    keep1 = rng.random(n) > 0.1
    keep2 = rng.random(n) > 0.1
    a = [(ids1[i], s1[i]) for i in range(n) if keep1[i]]
    b = [(ids2[i], s2[i]) for i in range(n) if keep2[i]]
    rng.shuffle(b)
    b1 = synth.to_fastq([x[1] for x in a], ids=[x[0] for x in a])
    b2 = synth.to_fastq([x[1] for x in b], ids=[x[0] for x in b])
    rc, r1, r2, so, _ = oracle.run_ref(tmp_path, "fast", oracle.FASTQ, b1, b2, unordered=True)
    o1, o2, st = oracle.run_oracle("fast", oracle.FASTQ, b1, b2, unordered=True)
    assert rc == 0 and o1 == r1 and o2 == r2
    assert _parse_counts(so)[:3] == [st.total, st.dups, st.unmatched]


def test_ref_error_paths(oracle, tmp_path):
    _need_ref(oracle)
    cases = [b"", b"@a\nACGT\n+\nFFF\n", b"@a\nACGT\n+\nFFFF\n@b\nACXT\n+\nFFFF\n@c\nAAAA\n+\nFFFF\n",
             b"@a\nACGT\n+\nFFFF\n@b\nAAAA\n+\nFFFF\nxc\nAAAA\n+\nFFFF\n", b"@a\nACGT\n+\nFFFF\n@b\nAAAA\n+\nFFFF",
             b"@a\nACGT\n+\nFFFF\n\n"]
    for k, buf in enumerate(cases):
        rc, r1, _, _, _ = oracle.run_ref(tmp_path / str(k), "fast", oracle.FASTQ, buf)
        out, _, st = oracle.run_oracle("fast", oracle.FASTQ, buf)
        assert (rc != 0) == (st.err != 0), (k, rc, st.err)
        assert out == r1, k


def _sweep_records(n, seed, mate=1, read_len=60):
    import synth
    buf = synth.to_fastq(synth.make_reads(n, seed=seed, read_len=read_len, dup_frac=0.3), mate=mate)
    lines = buf.split(b"\n")[:-1]
    return [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines), 4)]


def _sweep_damage(rec, kind):
    l = rec.split(b"\n")
    if kind == "start":
        l[0] = b"X" + l[0][1:]
    elif kind == "length":
        l[3] = l[3][:-3]
    else:
        l[1] = l[1][:10] + b"U" + l[1][11:]
    return b"\n".join(l)


@pytest.mark.parametrize("kind", ["start", "length", "base"])
def test_ref_malformed_record_at_every_position(oracle, tmp_path, kind):
    """One malformed record at every index, single- and paired-end (either mate): the oracle stops where the reference
    binary stops and has written the same bytes - including record 0, which the reference writes BEFORE keying it
    (src/hash_dup_remover.hpp:118-124,216-228), so a bad base there still leaves it in the output."""
    _need_ref(oracle)
    r1, r2 = _sweep_records(12, 50, 1), _sweep_records(12, 51, 2, read_len=45)
    for pos in range(12):
        bad1 = b"".join(r1[:pos] + [_sweep_damage(r1[pos], kind)] + r1[pos + 1:])
        bad2 = b"".join(r2[:pos] + [_sweep_damage(r2[pos], kind)] + r2[pos + 1:])
        good1, good2 = b"".join(r1), b"".join(r2)
        for k, (b1, b2) in enumerate([(bad1, None), (bad1, good2), (good1, bad2)]):
            rc, o1, o2, _, _ = oracle.run_ref(tmp_path / f"{pos}_{k}", "fast", oracle.FASTQ, b1, b2)
            e1, e2, st = oracle.run_oracle("fast", oracle.FASTQ, b1, b2)
            assert rc == 1 and st.err != 0, (pos, k)
            assert e1 == o1 and (b2 is None or e2 == o2), (pos, k)


@pytest.mark.parametrize("mode,unordered", [("tight", False), ("loose", False), ("tail-hamming", False), ("fast", True)])
def test_ref_malformed_record_in_whole_input_modes(oracle, tmp_path, mode, unordered):
    """Sequence-based modes and --fast --unordered read (and sort) everything before they write anything: a malformed
    record at any position, in either mate, ends the reference with exit status 1 and the outputs it has (or has not)
    created by then; the oracle reports an error and the same bytes."""
    _need_ref(oracle)
    r1, r2 = _sweep_records(12, 50, 1), _sweep_records(12, 51, 2, read_len=45)
    good1, good2 = b"".join(r1), b"".join(r2)
    for kind in ("start", "length", "base"):
        for pos in range(12):
            bad1 = b"".join(r1[:pos] + [_sweep_damage(r1[pos], kind)] + r1[pos + 1:])
            bad2 = b"".join(r2[:pos] + [_sweep_damage(r2[pos], kind)] + r2[pos + 1:])
            cases = [(bad1, good2), (good1, bad2)] + ([] if unordered else [(bad1, None)])
            for k, (b1, b2) in enumerate(cases):
                rc, o1, o2, _, _ = oracle.run_ref(tmp_path / f"{kind}_{pos}_{k}", mode, oracle.FASTQ, b1, b2,
                                                  stable=not unordered, unordered=unordered)
                e1, e2, st = oracle.run_oracle(mode, oracle.FASTQ, b1, b2, unordered=unordered) if unordered \
                    else oracle.run_oracle(mode, oracle.FASTQ, b1, b2)
                assert (rc != 0) == (st.err != 0), (kind, pos, k)
                assert e1 == (o1 or b"") and (b2 is None or (e2 or b"") == (o2 or b"")), (kind, pos, k)


def test_ref_unordered_many_tiny_cases(oracle, tmp_path):
    """300 random tiny --fast --unordered jobs against the reference binary: 0-8 records per file, tags drawn from a
    handful of values (many duplicate tags, missing mates, an empty tag), ID lines with the dot in different places,
    no dot, or no space - the cases where the end-of-stream rule of the two-pointer walk decides the output."""
    import random
    import shutil
    _need_ref(oracle)
    rng = random.Random(7)

    def ident(tag, mate, style):
        return [f"@R.{tag} {mate}", f"@R.{tag}", f"@{tag} x.{mate}", f"@R{tag}"][style].encode()

    def fastq(items):
        return b"".join(i + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in items)
    for it in range(300):
        n1, n2 = rng.randrange(0, 9), rng.randrange(0, 9)
        style = rng.choice([0, 0, 0, 1, 2, 3])
        tags = [str(rng.randrange(0, 6)) if rng.random() < 0.7 else rng.choice(["10", "2a", "A", "", "05"]) for _ in range(n1 + n2)]
        seqs = ["".join(rng.choice("ACGT") for _ in range(rng.choice([4, 4, 5]))).encode() for _ in range(4)]
        b1 = fastq([(ident(tags[i], 1, style), rng.choice(seqs)) for i in range(n1)])
        b2 = fastq([(ident(tags[n1 + i], 2, style), rng.choice(seqs)) for i in range(n2)])
        shutil.rmtree(tmp_path / "w", ignore_errors=True)
        rc, r1, r2, so, _ = oracle.run_ref(tmp_path / "w", "fast", oracle.FASTQ, b1, b2, unordered=True)
        o1, o2, st = oracle.run_oracle("fast", oracle.FASTQ, b1, b2, unordered=True)
        assert (rc != 0) == (st.err != 0), it
        assert (o1 or b"") == (r1 or b"") and (o2 or b"") == (r2 or b""), it
        if rc == 0:
            assert _parse_counts(so)[:3] == [st.total, st.dups, st.unmatched], it


def test_ref_stable_sequence_modes_many_tiny_cases(oracle, tmp_path):
    """300 random tiny jobs per run through every --compare-seq mode (distance 0-3, single- and paired-end) against the
    stable-sort build of the reference: 1-9 records of length 0-6 over {A,C,G,T,N}, made from three base sequences by
    cutting prefixes and substituting single bases - dense in exactly the relations the comparators look at."""
    import random
    import shutil
    _need_ref(oracle, stable=True)
    rng = random.Random(11)

    def fastq(seqs, mate):
        return b"".join(b"@r%d %d\n" % (i, mate) + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in enumerate(seqs))
    for it in range(300):
        k = rng.randrange(1, 10)
        base = ["".join(rng.choice("ACGTN" if rng.random() < 0.9 else "AC") for _ in range(rng.choice([0, 1, 2, 3, 3, 4, 4, 5, 6]))).encode()
                for _ in range(3)]

        def pick():
            s = rng.choice(base)
            r = rng.random()
            if r < 0.3 and s:
                s = s[:rng.randrange(0, len(s) + 1)]
            elif r < 0.5 and s:
                j = rng.randrange(len(s))
                s = s[:j] + bytes([rng.choice(b"ACGT")]) + s[j + 1:]
            return s
        s1 = [pick() for _ in range(k)]
        paired = rng.random() < 0.5
        s2 = [pick() for _ in range(k)] if paired else None
        mode, dist = rng.choice(["tight", "loose", "tail-hamming"]), rng.choice([0, 1, 2, 3])
        b1, b2 = fastq(s1, 1), (fastq(s2, 2) if paired else None)
        shutil.rmtree(tmp_path / "w", ignore_errors=True)
        rc, r1, r2, _, _ = oracle.run_ref(tmp_path / "w", mode, oracle.FASTQ, b1, b2, dist=dist, stable=True)
        o1, o2, st = oracle.run_oracle(mode, oracle.FASTQ, b1, b2, dist=dist)
        assert (rc != 0) == (st.err != 0), it
        assert (o1 or b"") == (r1 or b"") and (not paired or (o2 or b"") == (r2 or b"")), (it, mode, dist)


def test_ref_stable_cluster_files_many_tiny_cases(oracle, tmp_path):
    """150 tiny random jobs with --write-clusters in every --compare-seq mode: the text of <out>.clusters (both files in
    paired mode) equals what the stable-sort build of the reference writes."""
    import random
    import shutil
    _need_ref(oracle, stable=True)
    rng = random.Random(21)

    def fastq(seqs, mate):
        return b"".join(b"@r%d.%d %d\n" % (i, rng.randrange(9), mate) + s + b"\n+\n" + b"I" * len(s) + b"\n" for i, s in enumerate(seqs))
    for it in range(150):
        k = rng.randrange(1, 10)
        base = ["".join(rng.choice("ACGTN") for _ in range(rng.choice([1, 2, 3, 4, 5, 6]))).encode() for _ in range(3)]

        def pick():
            s = rng.choice(base)
            r = rng.random()
            if r < 0.3:
                s = s[:rng.randrange(1, len(s) + 1)]
            elif r < 0.5:
                j = rng.randrange(len(s))
                s = s[:j] + bytes([rng.choice(b"ACGT")]) + s[j + 1:]
            return s
        s1 = [pick() for _ in range(k)]
        paired = rng.random() < 0.5
        s2 = [pick() for _ in range(k)] if paired else None
        mode, dist = rng.choice(["tight", "loose", "tail-hamming"]), rng.choice([0, 1, 2])
        b1, b2 = fastq(s1, 1), (fastq(s2, 2) if paired else None)
        work = tmp_path / "w"
        shutil.rmtree(work, ignore_errors=True)
        rc, _, _, _, _ = oracle.run_ref(work, mode, oracle.FASTQ, b1, b2, dist=dist, stable=True, extra=["--write-clusters"])
        cl, _ = oracle.cluster_text(mode, oracle.FASTQ, b1, b2, dist=dist)
        assert rc == 0
        assert (work / "out_1.fq.clusters").read_bytes() == cl[0], (it, mode, dist)
        if paired:
            assert (work / "out_2.fq.clusters").read_bytes() == cl[1], (it, mode, dist)

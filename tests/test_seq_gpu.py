"""GPU parity tests for the sequence-based modes (sort + comparator scan) against the CPU oracle, whose sequential
scan and stable order are pinned to the reference (tests/test_oracle.py)."""
import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu

ERRMAP = {0: 0, 1: 3, 2: 4, 3: 5, 4: 6}


def _fx(golden_dir, kind, name):
    return (golden_dir / "ref_fixtures" / kind / name).read_bytes()


def _check(fqd, oracle, mode, b1, b2, fmt, dist=2, **kw):
    o1, o2, st = fqd.dedup_whole(mode, b1, b2, fmt, dist=dist, **kw)
    e1, e2, est = oracle.run_oracle(mode, fmt, b1, b2, dist=dist)
    assert st.err == ERRMAP[est.err]
    assert o1 == e1
    assert o2 == e2
    if est.err == 0:
        assert (st.total, st.dups) == (est.total, est.dups)
    return st


@pytest.mark.parametrize("name,mode,dist", [("single_tight.fa", "tight", 2), ("single_loose.fa", "loose", 2),
                                            ("single_hamming.fa", "tail-hamming", 1)])
def test_reference_fixtures_single(fqd, golden_dir, name, mode, dist):
    # test/test_seq.py:7-38
    out, _, _ = fqd.dedup_whole(mode, _fx(golden_dir, "inputs", name), None, fqd.FORMAT_FASTA, dist=dist)
    assert out == _fx(golden_dir, "expected", name)


def test_reference_fixture_paired_tight(fqd, golden_dir):
    # test/test_seq.py:41-75 - output is in SORTED order (00003, 00001, 00004)
    o1, o2, _ = fqd.dedup_whole("tight", _fx(golden_dir, "inputs", "paired_tight_r1.fa"), _fx(golden_dir, "inputs", "paired_tight_r2.fa"),
                                fqd.FORMAT_FASTA)
    assert o1 == _fx(golden_dir, "expected", "paired_tight_r1.fa")
    assert o2 == _fx(golden_dir, "expected", "paired_tight_r2.fa")


def test_negative_control(fqd, golden_dir):
    # test/test_seq.py:78-97
    out, _, _ = fqd.dedup_whole("tight", _fx(golden_dir, "inputs", "single_hamming.fa"), None, fqd.FORMAT_FASTA)
    assert out != _fx(golden_dir, "expected", "single_hamming.fa")


MODES = [("tight", 2), ("loose", 2), ("tail-hamming", 0), ("tail-hamming", 2), ("tail-hamming", 3)]


@pytest.mark.parametrize("mode,dist", MODES)
@pytest.mark.parametrize("fmt", ["fastq", "fasta"])
def test_random_single(fqd, oracle, mode, dist, fmt):
    kw = dict(read_len=60, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    seqs = synth.make_reads(6000, seed=31, **kw)
    f = fqd.FORMAT_FASTQ if fmt == "fastq" else fqd.FORMAT_FASTA
    buf = synth.to_fastq(seqs) if fmt == "fastq" else synth.to_fasta(seqs)
    _check(fqd, oracle, mode, buf, None, f, dist=dist, max_seq_len=60, seg_bytes=1 << 17, append_bytes=50_000)


@pytest.mark.parametrize("mode,dist", MODES)
def test_random_paired(fqd, oracle, mode, dist):
    kw = dict(read_len=45, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    s1, s2 = synth.make_pair(5000, seed=32, **kw)
    s2 = s2[:-9]
    b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    _check(fqd, oracle, mode, b1, b2, fqd.FORMAT_FASTQ, dist=dist, max_seq_len=45, seg_bytes=1 << 17)


def test_long_shared_prefixes_need_many_sort_rounds(fqd, oracle):
    # every read shares its first 45 bases (amplicon-like): ordering is decided by the 3rd key word and later
    rng = np.random.default_rng(33)
    prefix = bytes(rng.choice(list(b"ACGT"), size=45).astype(np.uint8))
    tails = synth.make_reads(4000, seed=34, read_len=70, var_len=True, dup_frac=0.4, prefix_frac=0.3, sub_frac=0.3)
    seqs = [prefix + t for t in tails]
    for mode, dist in (("tight", 2), ("loose", 2), ("tail-hamming", 2)):
        _check(fqd, oracle, mode, synth.to_fastq(seqs), None, fqd.FORMAT_FASTQ, dist=dist, max_seq_len=120)


@pytest.mark.parametrize("paired", [False, True])
def test_groups_of_every_size_behind_one_key_word(fqd, oracle, paired):
    """Groups of 2 .. 3000 reads share the first key word (21 bases) and differ - or not - further on: the sort hands
    groups up to 16 to one thread, up to 1024 to one block (rank among the members, ties by input order) and keeps only
    the larger ones for the next round; the order must stay the stable order of the reference's sort."""
    rng = np.random.default_rng(38)
    seqs = []
    for size in (2, 3, 15, 16, 17, 18, 40, 100, 255, 256, 257, 700, 1023, 1024, 1025, 3000):
        prefix = bytes(rng.choice(list(b"ACGT"), size=30).astype(np.uint8))
        variants = [bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(0, 60))).astype(np.uint8)) for _ in range(max(2, size // 3))]
        seqs += [prefix + variants[int(k)] for k in rng.integers(0, len(variants), size=size)]
    seqs += synth.make_reads(3000, seed=39, read_len=90, dup_frac=0.2)
    order = rng.permutation(len(seqs))
    seqs = [seqs[int(k)] for k in order]
    b1, b2 = synth.to_fastq(seqs, mate=1), None
    if paired:
        mates = [seqs[int(k)][:int(rng.integers(20, 90))] for k in rng.integers(0, len(seqs), size=len(seqs))]
        b2 = synth.to_fastq(mates, mate=2)
    for mode, dist in (("tight", 2), ("loose", 2), ("tail-hamming", 2)):
        _check(fqd, oracle, mode, b1, b2, fqd.FORMAT_FASTQ, dist=dist, max_seq_len=90, seg_bytes=1 << 20)


def test_150bp_with_exact_duplicates(fqd, oracle):
    seqs = synth.make_reads(20000, seed=35, read_len=150, dup_frac=0.3)
    st = _check(fqd, oracle, "tight", synth.to_fastq(seqs), None, fqd.FORMAT_FASTQ, max_seq_len=150, seg_bytes=1 << 20)
    assert 0.2 < st.dups / st.total < 0.4


def test_big_identical_cluster(fqd, oracle):
    # thousands of identical reads: resolved without refinement rounds (identical rows stay in index order)
    base = synth.make_reads(5, seed=36, read_len=100, dup_frac=0.0)
    rng = np.random.default_rng(37)
    seqs = [base[int(k)] for k in rng.integers(0, 5, size=20000)]
    st = _check(fqd, oracle, "tight", synth.to_fastq(seqs), None, fqd.FORMAT_FASTQ, max_seq_len=100)
    assert st.total - st.dups == 5
    _check(fqd, oracle, "tail-hamming", synth.to_fastq(seqs), None, fqd.FORMAT_FASTQ, max_seq_len=100)


def test_error_paths(fqd, oracle):
    for buf in (b"", b"@a\nACGT\n+\nFFF\n", b"@a\nACGT\n+\nFFFF\nxb\nACGT\n+\nFFFF\n", b"@a\nACGT\n+\nFFFF\n@b\nAC"):
        _check(fqd, oracle, "tight", buf, None, fqd.FORMAT_FASTQ)
    # bytes outside A/C/G/T/N are legal for the reference in this mode; this build reports them explicitly
    _, _, st = fqd.dedup_whole("tight", b"@a\nACGT\n+\nFFFF\n@b\nacgt\n+\nFFFF\n", None, fqd.FORMAT_FASTQ)
    assert st.err == 9


@pytest.mark.parametrize("mode,dist", [("tight", 2), ("loose", 2), ("tail-hamming", 2)])
@pytest.mark.parametrize("paired", [False, True])
def test_cluster_files(fqd, oracle, mode, dist, paired):
    """--write-clusters: fqd_emit_clusters against the oracle's cluster text (pinned to the reference in test_oracle.py);
    small emission buffer, several input segments."""
    kw = dict(read_len=40, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    if paired:
        s1, s2 = synth.make_pair(3000, seed=71, **kw)
        bufs = [synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)]
    else:
        bufs = [synth.to_fastq(synth.make_reads(4000, seed=72, **kw), id_fmt="@SYN.{i} some description {mate}")]
    texts, est = oracle.cluster_text(mode, oracle.FASTQ, bufs[0], bufs[1] if paired else None, dist=dist)
    eng = fqd.Engine(mode, fqd.FORMAT_FASTQ, paired, False, dist, 40, 5000, 1 << 17, 0, 0)
    try:
        for m, b in enumerate(bufs):
            for o in range(0, len(b), 70_000):
                eng.append(m, b[o: o + 70_000])
        eng.finish()
        st = eng.stats()
        assert st.err == 0 and (st.total, st.dups) == (est.total, est.dups)
        for m in range(len(bufs)):
            assert eng.emit_clusters_all(m, cap=3000) == texts[m]
    finally:
        eng.close()


@pytest.mark.parametrize("mode,dist", [("tight", 2), ("loose", 2), ("tail-hamming", 1), ("tail-hamming", 3)])
@pytest.mark.parametrize("paired", [False, True])
def test_arbitrary_bytes_with_byte_keys(fqd, oracle, mode, dist, paired):
    """Any byte is a legal sequence symbol in sequence-based modes (src/fastqview.cpp:56-67): with 3-bit rows the engine
    reports FQD_ERR_UNSUPPORTED_BYTE (the host then restarts with cfg.byte_keys), with byte rows it matches the oracle."""
    kw = dict(read_len=45, var_len=True, min_len=0, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5, alphabet=b"ACGTNacgtnRYKM*-.\t")
    if paired:
        s1, s2 = synth.make_pair(4000, seed=91, **kw)
        b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    else:
        b1, b2 = synth.to_fastq(synth.make_reads(6000, seed=92, **kw)), None
    _, _, st = fqd.dedup_whole(mode, b1, b2, fqd.FORMAT_FASTQ, dist=dist, max_seq_len=45, seg_bytes=1 << 17)
    assert st.err == 9        # FQD_ERR_UNSUPPORTED_BYTE
    _check(fqd, oracle, mode, b1, b2, fqd.FORMAT_FASTQ, dist=dist, max_seq_len=45, seg_bytes=1 << 17, append_bytes=60_000, byte_keys=1)


def test_byte_keys_on_plain_bases_match_code_keys(fqd, oracle):
    """Same input, both row formats: identical output (the order of {\\n,A,C,G,N,T} is the byte order)."""
    seqs = synth.make_reads(5000, seed=93, read_len=70, var_len=True, n_frac=0.1, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    buf = synth.to_fastq(seqs)
    for mode in ("tight", "loose", "tail-hamming"):
        a, _, sa = fqd.dedup_whole(mode, buf, None, fqd.FORMAT_FASTQ, max_seq_len=70, byte_keys=0)
        b, _, sb = fqd.dedup_whole(mode, buf, None, fqd.FORMAT_FASTQ, max_seq_len=70, byte_keys=1)
        assert a == b and (sa.total, sa.dups) == (sb.total, sb.dups)


@pytest.mark.parametrize("mode,dist,paired", [("tight", 2, False), ("loose", 2, True), ("tail-hamming", 2, True)])
@pytest.mark.parametrize("seg", [1 << 14, 1 << 16, 1 << 22])
def test_adopted_device_input(fqd, oracle, mode, dist, paired, seg):
    """fqd_adopt_device: the input is parsed in place through views cut at record boundaries (unaligned view starts,
    views smaller than the input, a trailing incomplete record) - same output as the copying path / the oracle."""
    kw = dict(read_len=70, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    if paired:
        s1, s2 = synth.make_pair(3000, seed=101, **kw)
        bufs = [synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2) + b"@incomplete\nACGT"]
    else:
        bufs = [synth.to_fastq(synth.make_reads(5000, seed=102, **kw))]
    e1, e2, est = oracle.run_oracle(mode, oracle.FASTQ, bufs[0], bufs[1] if paired else None, dist=dist)
    eng = fqd.Engine(mode, fqd.FORMAT_FASTQ, paired, False, dist, 70, 6000, seg, 0, 0)
    dbufs = []
    try:
        for m, b in enumerate(bufs):
            d = fqd.DeviceBuffer(len(b) + 64, 0)
            d.upload(b)
            dbufs.append(d)
            eng.adopt_device(m, d.ptr, len(b))
        eng.finish()
        st = eng.stats()
        assert st.err == 0 and (st.total, st.dups) == (est.total, est.dups)
        assert eng.emit_all(0, 1 << 15) == e1
        if paired:
            assert eng.emit_all(1, 1 << 15) == e2
    finally:
        eng.close()
        for d in dbufs:
            d.free()


def test_large_cross_mode_properties(fqd):
    """Size-independent properties at 3 M pairs (the oracle would need minutes): on a stream whose duplicates are exact
    copies of both mates, --compare-seq tight and --fast remove the same number of pairs (different representatives
    and output order, same count); loose and tail-hamming (d = 2) can only remove more on the same data; tight is
    idempotent (its own output holds no duplicates); the sorted output really is sorted."""
    lib = fqd.load_library()
    n, L = 3_000_000, 150
    rb = lib.fqd_synth_record_bytes(L)
    raw = [fqd.DeviceBuffer(n * rb + 4096) for _ in range(2)]
    for m in range(2):
        assert lib.fqd_synth_fastq(0, raw[m].ptr, 0, n, L, m + 1, 7, 300, 1, 0) == 0
    fast = fqd.Engine("fast", fqd.FORMAT_FASTQ, True, False, 2, L, n + 16, n * rb + 4096, n + 16, 0)
    res = fast.push_device(raw[0].ptr, n * rb, raw[1].ptr, n * rb)
    fast_dups = n - res.n_survivors
    fast.close()
    assert 0.25 < fast_dups / n < 0.35
    dups = {}
    for mode in ("tight", "loose", "tail-hamming"):
        eng = fqd.Engine(mode, fqd.FORMAT_FASTQ, True, False, 2, L, n + 16, 1 << 28, 0, 0)
        for m in range(2):
            eng.adopt_device(m, raw[m].ptr, n * rb)
        eng.finish()
        st = eng.stats()
        assert st.err == 0 and st.total == n
        dups[mode] = st.dups
        if mode == "tight":
            outs = [eng.emit_all(m, 1 << 26) for m in range(2)]
        eng.close()
    assert dups["tight"] == fast_dups
    assert dups["loose"] >= dups["tight"] and dups["tail-hamming"] >= dups["tight"]
    # the tight output: sorted by (R1 sequence, R2 sequence), no two equal neighbours; a second pass removes nothing
    seq1 = np.frombuffer(outs[0], dtype=np.uint8).reshape(-1, rb)[:, 18:18 + L]
    seq2 = np.frombuffer(outs[1], dtype=np.uint8).reshape(-1, rb)[:, 18:18 + L]
    assert seq1.shape[0] == n - dups["tight"]
    key = np.concatenate([seq1, seq2], axis=1)
    a, b = key[:-1], key[1:]
    neq = a != b
    first = neq.argmax(axis=1)
    assert neq.any(axis=1).all()                                   # no equal neighbours
    rows = np.arange(len(first))
    assert (a[rows, first] < b[rows, first]).all()                 # strictly increasing: byte order with N between G and T
    eng = fqd.Engine("tight", fqd.FORMAT_FASTQ, True, False, 2, L, n + 16, 1 << 28, 0, 0)
    eng.append(0, outs[0]); eng.append(1, outs[1])
    eng.finish()
    st = eng.stats()
    assert st.err == 0 and st.dups == 0 and st.total == n - dups["tight"]
    eng.close()
    for d in raw:
        d.free()


@pytest.mark.parametrize("mode,dist,paired,unordered", [("tight", 2, False, False), ("tail-hamming", 2, True, False), ("fast", 2, True, True)])
def test_record_tables_grow_in_place(fqd, oracle, mode, dist, paired, unordered):
    """max_records far below the input (a pipe has no size): the record tables and key rows grow while the input streams in."""
    kw = dict(read_len=50, var_len=True, min_len=5, n_frac=0.02, prefix_frac=0.2, sub_frac=0.2, dup_frac=0.4)
    if paired:
        s1, s2 = synth.make_pair(7000, seed=51, **kw)
        b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    else:
        b1, b2 = synth.to_fastq(synth.make_reads(9000, seed=52, **kw)), None
    if unordered:
        o1, o2, st = fqd.dedup_whole("fast", b1, b2, fqd.FORMAT_FASTQ, unordered=True, max_seq_len=50, seg_bytes=1 << 16, max_records=300)
        e1, e2, est = oracle.run_oracle("fast", oracle.FASTQ, b1, b2, unordered=True)
        assert (o1, o2) == (e1, e2) and (st.total, st.dups, st.unmatched) == (est.total, est.dups, est.unmatched)
    else:
        _check(fqd, oracle, mode, b1, b2, fqd.FORMAT_FASTQ, dist=dist, max_seq_len=50, seg_bytes=1 << 16, max_records=300)


@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("dist", [0, 1, 2])
def test_long_hamming_chains(fqd, oracle, paired, dist):
    """Amplicon-like input: thousands of reads within a few substitutions of each other form ONE segment of the
    tail-hamming scan (no definite break inside); the block-cooperative continuation of the greedy scan (k_ham_long) must
    keep exactly the heads the reference's sequential loop keeps."""
    rng = np.random.default_rng(91 + dist)
    base = [bytes(rng.choice(list(b"ACGT"), size=60).astype(np.uint8)) for _ in range(3)]

    def mutate(s, k):
        s = bytearray(s)
        for _ in range(k):
            s[int(rng.integers(40, 60))] = int(rng.choice(list(b"ACGT")))
        return bytes(s)
    seqs = [mutate(base[int(rng.integers(0, 3))], int(rng.integers(0, 3))) for _ in range(6000)]
    seqs += [bytes(rng.choice(list(b"ACGT"), size=60).astype(np.uint8)) for _ in range(500)]
    rng.shuffle(seqs)
    b1 = synth.to_fastq(seqs, mate=1)
    b2 = None
    if paired:
        seqs2 = [mutate(base[int(rng.integers(0, 3))], int(rng.integers(0, 2))) for _ in range(len(seqs))]
        b2 = synth.to_fastq(seqs2, mate=2)
    _check(fqd, oracle, "tail-hamming", b1, b2, fqd.FORMAT_FASTQ, dist=dist, max_seq_len=60, seg_bytes=1 << 20)


# ---- inputs larger than device memory: the raw bytes are freed as they are parsed (fqd_discard_input) ----------------
@pytest.mark.parametrize("mode,dist", MODES)
@pytest.mark.parametrize("paired", [False, True])
def test_discarded_input_matches_oracle(fqd, oracle, mode, dist, paired):
    """Same jobs as above with nothing but key rows and record tables resident: many small segments (each freed after its
    parse, tails carried), records fetched on the host from the windowed emission list."""
    kw = dict(read_len=60, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    if paired:
        s1, s2 = synth.make_pair(5000, seed=91, **kw)
        b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    else:
        b1, b2 = synth.to_fastq(synth.make_reads(7000, seed=92, **kw)), None
    _check(fqd, oracle, mode, b1, b2, fqd.FORMAT_FASTQ, dist=dist, max_seq_len=60, seg_bytes=1 << 16, append_bytes=37_000,
           discard=True, window=613)


def test_discarded_input_error_paths_and_refusals(fqd, oracle):
    for buf in (b"", b"@a\nACGT\n+\nFFF\n", b"@a\nACGT\n+\nFFFF\nxb\nACGT\n+\nFFFF\n", b"@a\nACGT\n+\nFFFF\n@b\nAC"):
        _check(fqd, oracle, "tight", buf, None, fqd.FORMAT_FASTQ, discard=True)
    eng = fqd.Engine("tight", fqd.FORMAT_FASTQ, False, False, 2, 40, 100, 1 << 16, 0, 0)
    try:
        eng.discard_input(True)
        eng.append(0, b"@a\nACGT\n+\nFFFF\n@b\nACGT\n+\nFFFF\n@c\nTTTT\n+\nFFFF\n")
        with pytest.raises(fqd.FqdError):
            eng.discard_input(False)                  # not after the first append
        eng.finish()
        st = eng.stats()
        assert (st.err, st.total, st.dups) == (0, 3, 1)
        with pytest.raises(fqd.FqdError):
            eng.emit_all(0)                           # the device has no bytes to gather from
        with pytest.raises(fqd.FqdError):
            eng.emit_clusters_all(0)
        with pytest.raises(fqd.FqdError):
            eng.emission_read(0, 1, 2)                # beyond the two written records
        off, ln = eng.emission_read(0, 0, 2)
        assert off.tolist() == [0, 30] and ln.tolist() == [15, 15]
        off, ln, head = eng.cluster_read(0, 0, 3)
        assert off.tolist() == [0, 15, 30] and ln.tolist() == [15, 15, 15] and head.tolist() == [1, 0, 1]
        eng.reset()                                   # the setting survives a reset
        eng.append(0, b"@a\nACGT\n+\nFFFF\n")
        eng.finish()
        with pytest.raises(fqd.FqdError):
            eng.emit_all(0)
    finally:
        eng.close()


@pytest.mark.parametrize("mode,dist", [("tight", 2), ("loose", 2), ("tail-hamming", 2)])
@pytest.mark.parametrize("paired", [False, True])
def test_cluster_files_from_discarded_input(fqd, oracle, mode, dist, paired):
    """--write-clusters when the device kept no raw bytes: the text rebuilt on the host from fqd_cluster_read windows
    must be the oracle's cluster text."""
    kw = dict(read_len=40, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    if paired:
        s1, s2 = synth.make_pair(3000, seed=73, **kw)
        bufs = [synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)]
    else:
        bufs = [synth.to_fastq(synth.make_reads(4000, seed=74, **kw), id_fmt="@SYN.{i} some description {mate}")]
    texts, est = oracle.cluster_text(mode, oracle.FASTQ, bufs[0], bufs[1] if paired else None, dist=dist)
    eng = fqd.Engine(mode, fqd.FORMAT_FASTQ, paired, False, dist, 40, 5000, 1 << 16, 0, 0)
    try:
        eng.discard_input(True)
        for m, b in enumerate(bufs):
            for o in range(0, len(b), 70_000):
                eng.append(m, b[o: o + 70_000])
        eng.finish()
        st = eng.stats()
        assert st.err == 0 and (st.total, st.dups) == (est.total, est.dups)
        for m, b in enumerate(bufs):
            text = []
            for k in range(0, st.total, 777):
                off, ln, head = eng.cluster_read(m, k, min(777, st.total - k))
                for o, l, hd in zip(off.tolist(), ln.tolist(), head.tolist()):
                    line = b[o: b.index(b"\n", o, o + l) + 1]
                    text.append(line if hd else b"--" + line)
            assert b"".join(text) == texts[m]
    finally:
        eng.close()

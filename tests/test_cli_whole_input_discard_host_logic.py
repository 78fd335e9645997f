"""The whole-input host suite once more with FQD_WHOLE_INPUT=discard: the device (here: the test double) keeps no raw
bytes, the binary maps a plain input or spools a .gz / a pipe into an unlinked temporary file, reads the (offset, length)
emission lists window by window and gathers the written records - and the cluster files - itself (host/replay.hpp,
dup_remover.cpp::gather_from_replay).  FAKE_FQD_REQUIRE_DISCARD makes the double refuse fqd_emit, so a run that silently
took the resident path fails.  Then the cases of its own: the automatic choice by device memory, spool directory
problems, outputs that are pipes or .gz."""
import gzip
import os
import subprocess
import threading

import pytest

import synth
import test_cli_whole_input_host_logic as whole_suite
from test_cli_whole_input_host_logic import run, fake_engine  # noqa: F401  (module fixture builds the double + the binary)
from test_host_io import deflate_gz


@pytest.fixture(autouse=True)
def discarded_input(monkeypatch):
    monkeypatch.setenv("FQD_WHOLE_INPUT", "discard")
    monkeypatch.setenv("FAKE_FQD_REQUIRE_DISCARD", "1")


@pytest.mark.parametrize("mode,dist", whole_suite.MODES)
@pytest.mark.parametrize("paired", [False, True])
def test_sequence_modes(tmp_path, oracle, mode, dist, paired):
    # mate 1 is a multi-member .gz (spooled), mate 2 a plain file (mapped); .gz and plain outputs; cluster files
    whole_suite.test_sequence_modes_against_oracle_and_stable_reference(tmp_path, oracle, mode, dist, paired)


def test_unordered_with_restart(tmp_path, oracle):
    whole_suite.test_unordered_with_long_tags_restarts_and_matches_the_reference(tmp_path, oracle)


@pytest.mark.parametrize("mode", ["tight", "loose", "tail-hamming"])
def test_byte_key_restart(tmp_path, oracle, mode):
    whole_suite.test_arbitrary_sequence_bytes_restart_with_byte_keys(tmp_path, oracle, mode)


def test_capacity_and_row_width_restarts(tmp_path, oracle):
    whole_suite.test_restarts_on_capacity_and_row_width(tmp_path, oracle)


@pytest.mark.parametrize("mode,unordered", [("tight", False), ("tail-hamming", False), ("fast", True)])
def test_malformed_records(tmp_path, oracle, mode, unordered):
    whole_suite.test_malformed_record_in_whole_input_modes_matches_the_reference_binary(tmp_path, oracle, mode, unordered)


def test_odd_inputs(tmp_path, oracle):
    whole_suite.test_odd_inputs_in_every_mode_match_the_reference_binaries(tmp_path, oracle)


def test_unusual_files(tmp_path, oracle):
    whole_suite.test_unusual_files_in_every_mode_match_the_reference_binaries(tmp_path, oracle)


def test_fifo_in_fifo_out(tmp_path, oracle):
    whole_suite.test_fifo_input_and_fifo_output_in_a_sequence_mode(tmp_path, oracle)


# ---- cases of this path's own --------------------------------------------------------------------------------------
def _job(tmp_path, n=20000, seed=5):
    s1, s2 = synth.make_pair(n, seed=seed, read_len=50, var_len=True, prefix_frac=0.2, sub_frac=0.2, dup_frac=0.4)
    b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    (tmp_path / "a.fq").write_bytes(b1)
    (tmp_path / "b.fq").write_bytes(b2)
    return b1, b2


def test_choice_follows_device_memory(tmp_path, oracle, monkeypatch):
    """Without FQD_WHOLE_INPUT the host compares the input's size with the device's free memory: a device with room keeps
    the raw bytes (fqd_emit gathers), a small one makes the host gather; both give the oracle's bytes."""
    monkeypatch.delenv("FQD_WHOLE_INPUT")
    monkeypatch.delenv("FAKE_FQD_REQUIRE_DISCARD")
    b1, b2 = _job(tmp_path)
    (tmp_path / "a.fq.gz").write_bytes(deflate_gz(b1, 6))       # a .gz would have to be spooled: it stays on a device with room
    e1, e2, est = oracle.run_oracle("loose", oracle.FASTQ, b1, b2)
    io = ["-i", tmp_path / "a.fq.gz", "-u", tmp_path / "b.fq", "-o", tmp_path / "o1.fq", "-p", tmp_path / "o2.fq", "--compare-seq", "loose", "-v"]
    roomy = run(*io, env={"FQD_TRACE": "1"})
    assert roomy.returncode == 0, roomy.stderr
    assert "raw input not kept" not in roomy.stderr
    assert (tmp_path / "o1.fq").read_bytes() == e1 and (tmp_path / "o2.fq").read_bytes() == e2
    os.remove(tmp_path / "o1.fq"), os.remove(tmp_path / "o2.fq")
    tight = run(*io, env={"FQD_TRACE": "1", "FAKE_FQD_DEVICE_BYTES": str(2 << 30), "FAKE_FQD_REQUIRE_DISCARD": "1"})
    assert tight.returncode == 0, tight.stderr
    assert "raw input not kept" in tight.stderr and "gathered on the host" in tight.stderr
    assert (tmp_path / "o1.fq").read_bytes() == e1 and (tmp_path / "o2.fq").read_bytes() == e2
    assert tight.stdout == roomy.stdout == f"{est.total} read pairs processed, out of which {est.dups} duplicates were removed.\n"
    # the resident path can be forced as well
    forced = run(*io, env={"FQD_TRACE": "1", "FAKE_FQD_DEVICE_BYTES": str(2 << 30), "FQD_WHOLE_INPUT": "resident"})
    assert forced.returncode == 0 and "raw input not kept" not in forced.stderr
    # plain files that fit the page cache are mapped and gathered on the host even when the device has room
    io_plain = ["-i", tmp_path / "a.fq", "-u", tmp_path / "b.fq", "-o", tmp_path / "p1.fq", "-p", tmp_path / "p2.fq", "--compare-seq", "loose", "-v"]
    plain = run(*io_plain, env={"FQD_TRACE": "1", "FAKE_FQD_REQUIRE_DISCARD": "1"})
    assert plain.returncode == 0, plain.stderr
    assert "raw input not kept" in plain.stderr and (tmp_path / "p1.fq").read_bytes() == e1 and (tmp_path / "p2.fq").read_bytes() == e2


def test_spool_goes_where_it_is_told_and_leaves_nothing_behind(tmp_path, oracle):
    b1, _ = _job(tmp_path, n=8000)
    (tmp_path / "a.fq.gz").write_bytes(deflate_gz(b1, 6))
    spool = tmp_path / "spool"
    spool.mkdir()
    out = tmp_path / "out"
    out.mkdir()
    res = run("-i", tmp_path / "a.fq.gz", "-o", out / "o.fq", "--compare-seq", "tight", env={"FQD_SPOOL_DIR": str(spool)})
    assert res.returncode == 0, res.stderr
    e1, _, _ = oracle.run_oracle("tight", oracle.FASTQ, b1)
    assert (out / "o.fq").read_bytes() == e1
    assert list(spool.iterdir()) == [] and [p.name for p in out.iterdir()] == ["o.fq"]
    # default: next to the output file; unlinked at once
    res = run("-i", tmp_path / "a.fq.gz", "-o", out / "o2.fq", "--compare-seq", "tight")
    assert res.returncode == 0 and sorted(p.name for p in out.iterdir()) == ["o.fq", "o2.fq"]
    # a spool directory that does not exist is an error with advice, not a crash
    res = run("-i", tmp_path / "a.fq.gz", "-o", out / "o3.fq", "--compare-seq", "tight", env={"FQD_SPOOL_DIR": str(tmp_path / "missing")})
    assert res.returncode != 0 and "input spool" in res.stderr and "FQD_SPOOL_DIR" in res.stderr


def test_gz_and_pipe_outputs(tmp_path, oracle):
    b1, b2 = _job(tmp_path, n=12000, seed=6)
    e1, e2, _ = oracle.run_oracle("tail-hamming", oracle.FASTQ, b1, b2, dist=2)
    os.mkfifo(tmp_path / "out.fifo")
    got = {}

    def drain():
        with open(tmp_path / "out.fifo", "rb") as f:
            got["out"] = f.read()
    t = threading.Thread(target=drain)
    t.start()
    res = run("-i", tmp_path / "a.fq", "-u", tmp_path / "b.fq", "-o", tmp_path / "o1.fq.gz", "-p", tmp_path / "out.fifo",
              "--compare-seq", "tail-hamming", "--distance", 2)
    t.join(timeout=60)
    assert res.returncode == 0, res.stderr
    assert gzip.decompress((tmp_path / "o1.fq.gz").read_bytes()) == e1
    assert got["out"] == e2


def test_many_list_windows(tmp_path, oracle):
    """More written records than one list window (FQD_LIST_WINDOW shrinks the 2^20 of production): dozens of windows, three
    in flight behind the asynchronous writer, records and cluster lines gathered window by window."""
    seqs = synth.make_reads(40_000, seed=8, read_len=24, dup_frac=0.1)
    buf = synth.to_fasta(seqs)
    (tmp_path / "a.fa").write_bytes(buf)
    res = run("-i", tmp_path / "a.fa", "-o", tmp_path / "o.fa", "--format", "fasta", "--compare-seq", "tight", "-v", "--write-clusters",
              env={"FQD_LIST_WINDOW": "777"})
    assert res.returncode == 0, res.stderr
    e1, _, est = oracle.run_oracle("tight", oracle.FASTA, buf)
    assert est.total - est.dups > 30 * 777
    assert (tmp_path / "o.fa").read_bytes() == e1
    cl, _ = oracle.cluster_text("tight", oracle.FASTA, buf)
    assert (tmp_path / "o.fa.clusters").read_bytes() == cl[0]


def _feed(path, data):
    def go():
        with open(path, "wb") as f:
            f.write(data)
    t = threading.Thread(target=go)
    t.start()
    return t


@pytest.mark.parametrize("mode", ["tight", "tail-hamming"])
def test_pipe_input_survives_a_restart(tmp_path, oracle, monkeypatch, mode):
    """A FIFO cannot be read twice, its spool can: lower-case bases force the byte-key restart, a late long sequence the
    wider rows - both read the pipe's bytes from /proc/self/fd/<spool> the second time.  No FQD_WHOLE_INPUT: a pipe is
    spooled by default."""
    import random
    monkeypatch.delenv("FQD_WHOLE_INPUT")
    rng = random.Random(11)
    seqs = [bytes(rng.choice(b"ACGTNacgt") for _ in range(rng.choice([20, 25]))) for _ in range(60)]
    reads = [rng.choice(seqs) for _ in range(3000)] + [bytes(rng.choice(b"ACGT") for _ in range(130))] * 3
    buf = synth.to_fastq(reads)
    os.mkfifo(tmp_path / "in.fifo")
    t = _feed(tmp_path / "in.fifo", buf)
    res = run("-i", tmp_path / "in.fifo", "-o", tmp_path / "o.fq", "--compare-seq", mode, "-v", "--write-clusters",
              env={"FQD_BLOCK_BYTES": str(1 << 14), "FQD_TRACE": "1"})
    t.join(timeout=60)
    assert res.returncode == 0, res.stderr
    assert res.stderr.count("raw input not kept") >= 2          # at least one restart happened, on the spool
    e1, _, est = oracle.run_oracle(mode, oracle.FASTQ, buf)
    assert (tmp_path / "o.fq").read_bytes() == e1
    cl, _ = oracle.cluster_text(mode, oracle.FASTQ, buf)
    assert (tmp_path / "o.fq.clusters").read_bytes() == cl[0]
    assert res.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"
    assert sorted(p.name for p in tmp_path.iterdir()) == ["in.fifo", "o.fq", "o.fq.clusters"]      # the spool left nothing behind


def test_pipe_input_length_mismatch_quotes_the_record(tmp_path, oracle, monkeypatch):
    """The reference's message for len(seq) != len(qual) quotes both strings; with a pipe the record is fetched from the spool."""
    if not oracle.ref_available(stable=True):
        pytest.skip("oracle/_ref not built")
    monkeypatch.delenv("FQD_WHOLE_INPUT")
    from test_cli_host_logic import _damage, _records
    recs = _records(40, seed=12, mate=1)
    recs[23] = _damage(recs[23], "length")
    buf = b"".join(recs)
    (tmp_path / "a.fq").write_bytes(buf)
    ref = subprocess.run([str(oracle.REF_STABLE_BIN), "-i", "a.fq", "-o", "r.fq", "--compare-seq", "tight", "-v"], capture_output=True, text=True, cwd=tmp_path)
    os.mkfifo(tmp_path / "in.fifo")
    t = _feed(tmp_path / "in.fifo", buf)
    res = run("-i", tmp_path / "in.fifo", "-o", tmp_path / "o.fq", "--compare-seq", "tight", "-v", env={"FQD_BLOCK_BYTES": "4096"})
    t.join(timeout=60)
    assert (res.returncode, res.stdout, res.stderr) == (ref.returncode, ref.stdout, ref.stderr)
    assert (tmp_path / "o.fq").exists() == (tmp_path / "r.fq").exists()


def test_many_tiny_jobs_match_the_reference_binary(tmp_path, oracle):
    """200 random tiny jobs through the binary on the discarded-input path - every --compare-seq mode and --fast --unordered,
    single- and paired-end, FASTQ and FASTA, inputs plain (mapped) or .gz (spooled) by turns, cluster files - against the
    reference binaries: exit status, -v line, output bytes, <out>.clusters.  Records of length 0-6 built from three base
    sequences by cutting prefixes and substituting bases: dense in the relations the comparators look at, and in ties."""
    import random
    import shutil
    if not oracle.ref_available(stable=True):
        pytest.skip("oracle/_ref not built")
    rng = random.Random(23)

    def text(seqs, mate, fasta, tags=None):
        ids = tags if tags is not None else [b"r%d" % i for i in range(len(seqs))]
        if fasta:
            return b"".join(b">RUN." + t + b" %d\n" % mate + s + b"\n" for t, s in zip(ids, seqs))
        return b"".join(b"@RUN." + t + b" %d\n" % mate + s + b"\n+\n" + b"I" * len(s) + b"\n" for t, s in zip(ids, seqs))
    for it in range(200):
        k = rng.randrange(1, 10)
        base = ["".join(rng.choice("ACGTN") for _ in range(rng.choice([0, 1, 2, 3, 3, 4, 4, 5, 6]))).encode() for _ in range(3)]

        def pick():
            s = rng.choice(base)
            r = rng.random()
            if r < 0.3 and s:
                s = s[:rng.randrange(0, len(s) + 1)]
            elif r < 0.5 and s:
                j = rng.randrange(len(s))
                s = s[:j] + bytes([rng.choice(b"ACGT")]) + s[j + 1:]
            return s
        unordered = rng.random() < 0.25
        paired = unordered or rng.random() < 0.5
        fasta = rng.random() < 0.3
        s1 = [pick() for _ in range(k)]
        s2 = [pick() for _ in range(k)] if paired else None
        t1 = t2 = None
        if unordered:                    # tags from a small pool, both files sorted differently, some without a partner
            t1 = [b"%02d" % rng.randrange(12) for _ in range(k)]
            t2 = [b"%02d" % rng.randrange(12) for _ in range(k)]
        d = tmp_path / "w"
        shutil.rmtree(d, ignore_errors=True)
        d.mkdir()
        b1 = text(s1, 1, fasta, t1)
        b2 = text(s2, 2, fasta, t2) if paired else None
        gz = [rng.random() < 0.4, rng.random() < 0.4]
        names = ["a.gz" if gz[0] else "a", "b.gz" if gz[1] else "b"]
        (d / "a").write_bytes(b1)
        (d / names[0]).write_bytes(gzip.compress(b1) if gz[0] else b1)
        if paired:
            (d / "b").write_bytes(b2)
            (d / names[1]).write_bytes(gzip.compress(b2) if gz[1] else b2)
        if unordered:
            flags, ref_bin, clusters = ["--fast", "--unordered"], oracle.REF_BIN, False
        else:
            mode = rng.choice(["tight", "loose", "tail-hamming"])
            clusters = rng.random() < 0.5
            flags = ["--compare-seq", mode, "--distance", str(rng.choice([0, 1, 2, 3]))] + (["--write-clusters"] if clusters else [])
            ref_bin = oracle.REF_STABLE_BIN
        common = ["-v", "--format", "fasta" if fasta else "fastq", *flags]
        io_r = ["-i", "a", "-o", "r1"] + (["-u", "b", "-p", "r2"] if paired else [])
        io_o = ["-i", d / names[0], "-o", d / "o1"] + (["-u", d / names[1], "-p", d / "o2"] if paired else [])
        ref = subprocess.run([str(ref_bin), *io_r, *common], capture_output=True, text=True, cwd=d)
        ours = run(*io_o, *common, env={"FQD_BLOCK_BYTES": "4096"})
        where = (it, flags, paired, fasta, gz)
        assert (ours.returncode, ours.stdout, ours.stderr) == (ref.returncode, ref.stdout, ref.stderr), where
        for mine, theirs in (("o1", "r1"), ("o2", "r2"), ("o1.clusters", "r1.clusters"), ("o2.clusters", "r2.clusters")):
            assert (d / mine).exists() == (d / theirs).exists(), (where, mine)
            if (d / mine).exists():
                assert (d / mine).read_bytes() == (d / theirs).read_bytes(), (where, mine)

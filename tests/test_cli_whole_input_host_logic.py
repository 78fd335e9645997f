"""Sequence-based modes, --fast --unordered and --write-clusters through the real binary on the CPU: the test double
of the C ABI collects the appended input and lets the oracle's C functions decide (tests/fake_engine), so what is
exercised here is the driver around the engine - blocks streamed to fqd_append, restarts (byte keys for arbitrary
sequence bytes, wider tag rows, wider key rows, capacity), output pulled through two staging buffers behind an
asynchronous writer, cluster files, -v lines, error wording - compared with the oracle AND with the reference binaries
(the stable-sort build where the choice inside a tie group matters)."""
import gzip
import os
import subprocess
import sys
from pathlib import Path

import pytest

import synth
from test_host_io import SMALL, deflate_gz, members

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "fastq-dupaway_b200" / "host" / "fastq-dupaway"
sys.path.insert(0, str(ROOT / "tests" / "fake_engine"))
from build import BUILD as FAKE_DIR, build_fake  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def fake_engine():
    build_fake()
    if not EXE.exists():
        subprocess.run(["make", "-s", "-C", str(EXE.parent)], check=True)


def run(*args, env=None):
    e = dict(os.environ, LD_LIBRARY_PATH=str(FAKE_DIR), FQD_IO_THREADS="4")
    e.update(env or {})
    p = subprocess.run([str(EXE), *map(str, args)], capture_output=True, text=True, env=e, timeout=300)
    p.stderr = "".join(l + "\n" for l in p.stderr.splitlines() if not l.startswith("[fake_fqd]"))
    return p


MODES = [("tight", 2), ("loose", 2), ("tail-hamming", 0), ("tail-hamming", 3)]
THIN = False        # tests/test_differential_gpu.py: fewer positions / flag sets per test (every run pays a CUDA start-up)


@pytest.mark.parametrize("mode,dist", MODES)
@pytest.mark.parametrize("paired", [False, True])
def test_sequence_modes_against_oracle_and_stable_reference(tmp_path, oracle, mode, dist, paired):
    s1, s2 = synth.make_pair(6000, seed=70, read_len=60, var_len=True, prefix_frac=0.2, sub_frac=0.2, dup_frac=0.4, n_frac=0.02)
    b1, b2 = synth.to_fastq(s1, mate=1), (synth.to_fastq(s2, mate=2) if paired else None)
    (tmp_path / "a.fq.gz").write_bytes(members(b1, [300_000]))
    io = ["-i", tmp_path / "a.fq.gz", "-o", tmp_path / "o1.fq.gz"]
    if paired:
        (tmp_path / "b.fq").write_bytes(b2)
        io += ["-u", tmp_path / "b.fq", "-p", tmp_path / "o2.fq"]
    res = run(*io, "--compare-seq", mode, "--distance", dist, "-v", "--write-clusters", env={"FQD_BLOCK_BYTES": str(1 << 16)})
    assert res.returncode == 0, res.stderr
    e1, e2, est = oracle.run_oracle(mode, oracle.FASTQ, b1, b2, dist=dist)
    assert gzip.decompress((tmp_path / "o1.fq.gz").read_bytes()) == e1
    what = "read pairs" if paired else "reads"
    assert res.stdout == f"{est.total} {what} processed, out of which {est.dups} duplicates were removed.\n"
    cl, _ = oracle.cluster_text(mode, oracle.FASTQ, b1, b2, dist=dist)
    assert (tmp_path / "o1.fq.gz.clusters").read_bytes() == cl[0]
    if paired:
        assert (tmp_path / "o2.fq").read_bytes() == e2
        assert (tmp_path / "o2.fq.clusters").read_bytes() == cl[1]
    if oracle.ref_available(stable=True):
        rc, r1, r2, so, _ = oracle.run_ref(tmp_path / "ref", mode, oracle.FASTQ, b1, b2, dist=dist, stable=True)
        assert rc == 0 and r1 == e1 and (not paired or r2 == e2) and so == res.stdout


def test_unordered_with_long_tags_restarts_and_matches_the_reference(tmp_path, oracle):
    import numpy as np
    rng = np.random.default_rng(3)
    n = 3000
    s1, s2 = synth.make_pair(n, seed=71, read_len=40)
    long_tag = "x" * 50                                       # longer than the 32 bytes the first attempt allows
    ids1 = [f"@RUN.{long_tag}{i:06d} 1".encode() for i in range(n)]
    ids2 = [f"@RUN.{long_tag}{i:06d} 2".encode() for i in range(n)]
    keep1, keep2 = rng.random(n) > 0.1, rng.random(n) > 0.1
    a = [(ids1[i], s1[i]) for i in range(n) if keep1[i]]
    b = [(ids2[i], s2[i]) for i in range(n) if keep2[i]]
    rng.shuffle(b)
    b1 = synth.to_fastq([x[1] for x in a], ids=[x[0] for x in a])
    b2 = synth.to_fastq([x[1] for x in b], ids=[x[0] for x in b])
    (tmp_path / "a.fq").write_bytes(b1)
    (tmp_path / "b.fq.gz").write_bytes(deflate_gz(b2, 6))
    res = run("-i", tmp_path / "a.fq", "-u", tmp_path / "b.fq.gz", "-o", tmp_path / "o1.fq", "-p", tmp_path / "o2.fq",
              "--fast", "--unordered", "-v", env=dict(SMALL, FQD_BLOCK_BYTES=str(1 << 16), FQD_TRACE="1"))
    assert res.returncode == 0, res.stderr
    e1, e2, est = oracle.run_oracle("fast", oracle.FASTQ, b1, b2, unordered=True)
    assert (tmp_path / "o1.fq").read_bytes() == e1 and (tmp_path / "o2.fq").read_bytes() == e2
    assert res.stdout == (f"{est.total} valid read pairs processed, out of which {est.dups} duplicates were removed.\n"
                          f"{est.unmatched} Non-matching entries from both files were skipped.\n")
    if oracle.ref_available():
        rc, r1, r2, so, _ = oracle.run_ref(tmp_path / "ref", "fast", oracle.FASTQ, b1, b2, unordered=True)
        assert rc == 0 and r1 == e1 and r2 == e2 and so == res.stdout


@pytest.mark.parametrize("mode", ["tight", "loose", "tail-hamming"])
def test_arbitrary_sequence_bytes_restart_with_byte_keys(tmp_path, oracle, mode):
    import random
    rng = random.Random(72)
    alphabet = b"ACGTNacgtRYKM*-"
    seqs = [bytes(rng.choice(alphabet) for _ in range(rng.choice([20, 20, 25]))) for _ in range(40)]
    reads = [rng.choice(seqs) for _ in range(2000)]
    b1 = synth.to_fastq(reads)
    (tmp_path / "a.fq").write_bytes(b1)
    res = run("-i", tmp_path / "a.fq", "-o", tmp_path / "o.fq", "--compare-seq", mode, "-v", env={"FQD_BLOCK_BYTES": str(1 << 14)})
    assert res.returncode == 0, res.stderr
    e1, _, est = oracle.run_oracle(mode, oracle.FASTQ, b1)
    assert (tmp_path / "o.fq").read_bytes() == e1
    assert res.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"


def test_restarts_on_capacity_and_row_width(tmp_path, oracle):
    short = synth.make_reads(2000, seed=73, read_len=30, dup_frac=0.3)
    long_ = synth.make_reads(2000, seed=74, read_len=140, dup_frac=0.3)
    b1 = synth.to_fastq(short + long_)
    (tmp_path / "a.fq").write_bytes(b1)
    res = run("-i", tmp_path / "a.fq", "-o", tmp_path / "o.fq", "--compare-seq", "tight", "-v",
              env={"FQD_BLOCK_BYTES": str(1 << 14), "FAKE_FQD_SHRINK": "40"})
    assert res.returncode == 0, res.stderr
    e1, _, est = oracle.run_oracle("tight", oracle.FASTQ, b1)
    assert (tmp_path / "o.fq").read_bytes() == e1


@pytest.mark.parametrize("mode,unordered", [("tight", False), ("tail-hamming", False), ("fast", True)])
def test_malformed_record_in_whole_input_modes_matches_the_reference_binary(tmp_path, oracle, mode, unordered):
    """Exit status, stderr and which output files exist (with what in them) when one record is malformed, at a few
    positions, in either mate - the reference reads and sorts everything before it creates an output."""
    if not oracle.ref_available(stable=True):
        pytest.skip("oracle/_ref not built")
    from test_cli_host_logic import _damage, _records
    r = [_records(16, seed=75, mate=1), _records(16, seed=76, mate=2, read_len=80)]
    for kind in ("start", "length", "base"):
        for pos in ((0, 1, 7, 15) if not THIN else (7,)):
            for bad_mate in (0, 1):
                recs = [list(r[0]), list(r[1])]
                recs[bad_mate][pos] = _damage(r[bad_mate][pos], kind)
                work = tmp_path / f"{kind}_{pos}_{bad_mate}"
                work.mkdir()
                (work / "a.fq").write_bytes(b"".join(recs[0]))
                (work / "b.fq").write_bytes(b"".join(recs[1]))
                flags = ["--fast", "--unordered"] if unordered else ["--compare-seq", mode]
                ref_bin = oracle.REF_BIN if unordered else oracle.REF_STABLE_BIN
                ref = subprocess.run([str(ref_bin), "-i", "a.fq", "-u", "b.fq", "-o", "r1.fq", "-p", "r2.fq", "-v", *flags],
                                     capture_output=True, text=True, cwd=work)
                ours = run("-i", work / "a.fq", "-u", work / "b.fq", "-o", work / "o1.fq", "-p", work / "o2.fq", "-v", *flags,
                           env={"FQD_BLOCK_BYTES": "4096"})
                where = (kind, pos, bad_mate)
                assert ours.returncode == ref.returncode, where
                assert ours.stderr == ref.stderr, where
                assert ours.stdout == ref.stdout, where
                for mine, theirs in (("o1.fq", "r1.fq"), ("o2.fq", "r2.fq")):
                    assert (work / mine).exists() == (work / theirs).exists(), where
                    if (work / mine).exists():
                        assert (work / mine).read_bytes() == (work / theirs).read_bytes(), where


def test_odd_inputs_in_every_mode_match_the_reference_binaries(tmp_path, oracle):
    """Files of different length, an empty mate, an empty file, one record, no final newline, a blank line at the end,
    FASTA pairs - through tight, loose and --fast [--unordered]: exit status, stdout, stderr, which outputs exist and
    their bytes are the reference binaries'."""
    import shutil
    if not oracle.ref_available(stable=True):
        pytest.skip("oracle/_ref not built")
    s1, s2 = synth.make_pair(60, seed=80, read_len=40, dup_frac=0.4)
    cases = {
        "pe_r2_shorter": (synth.to_fastq(s1, mate=1), synth.to_fastq(s2[:45], mate=2), "fastq"),
        "pe_r1_shorter": (synth.to_fastq(s1[:45], mate=1), synth.to_fastq(s2, mate=2), "fastq"),
        "pe_r2_empty": (synth.to_fastq(s1, mate=1), b"", "fastq"),
        "se_empty": (b"", None, "fastq"),
        "fasta_pe": (synth.to_fasta(s1, mate=1), synth.to_fasta(s2, mate=2), "fasta"),
        "se_one_record": (synth.to_fastq(s1[:1]), None, "fastq"),
        "se_no_final_newline": (synth.to_fastq(s1)[:-1], None, "fastq"),
        "se_blank_line_end": (synth.to_fastq(s1) + b"\n", None, "fastq"),
    }
    for name, (b1, b2, fmt) in cases.items():
        for flags in (["--compare-seq", "tight"], ["--compare-seq", "loose"], ["--fast"] + (["--unordered"] if b2 is not None else []))[1 if THIN else 0:]:
            d = tmp_path / name
            shutil.rmtree(d, ignore_errors=True)
            d.mkdir()
            (d / "a").write_bytes(b1)
            io_r, io_o = ["-i", "a", "-o", "r1"], ["-i", d / "a", "-o", d / "o1"]
            if b2 is not None:
                (d / "b").write_bytes(b2)
                io_r += ["-u", "b", "-p", "r2"]
                io_o += ["-u", d / "b", "-p", d / "o2"]
            common = ["-v", "--format", fmt, *flags]
            ref_bin = oracle.REF_STABLE_BIN if flags[0] == "--compare-seq" else oracle.REF_BIN
            ref = subprocess.run([str(ref_bin), *io_r, *common], capture_output=True, text=True, cwd=d)
            ours = run(*io_o, *common, env={"FQD_BLOCK_BYTES": "4096"})
            where = (name, flags)
            assert (ours.returncode, ours.stdout, ours.stderr) == (ref.returncode, ref.stdout, ref.stderr), where
            for mine, theirs in (("o1", "r1"), ("o2", "r2")):
                assert (d / mine).exists() == (d / theirs).exists(), where
                if (d / mine).exists():
                    assert (d / mine).read_bytes() == (d / theirs).read_bytes(), where


def test_the_references_own_test_suite_passes_against_the_binary(tmp_path):
    """`pytest <reference>/test`, unchanged, from the directory of the drop-in binary (it looks for ./fastq-dupaway):
    its 14 tests (help, fast SE/PE, every sequence-based mode, unordered) pass.  Only where the reference tree is
    mounted; the same fixtures run against the real engine in tests/test_cli_gpu.py."""
    ref_tests = Path("/root/reference/test")
    if not ref_tests.is_dir():
        pytest.skip("reference tree not present")
    e = dict(os.environ, LD_LIBRARY_PATH=str(FAKE_DIR), FQD_IO_THREADS="4")
    p = subprocess.run([sys.executable, "-m", "pytest", str(ref_tests), "-q", "-p", "no:cacheprovider"], capture_output=True, text=True,
                       cwd=EXE.parent, env=e, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]
    assert " passed" in p.stdout and "failed" not in p.stdout


def test_unusual_files_in_every_mode_match_the_reference_binaries(tmp_path, oracle):
    """CRLF line ends, multi-line FASTA, the wrong --format, tabs and '@' where they do not belong, empty sequence lines,
    '+' lines that repeat the ID, binary garbage, a file of newlines - through --fast and every --compare-seq mode:
    exit status, stdout, stderr, whether an output exists and its bytes are the reference binaries'."""
    import shutil
    if not oracle.ref_available(stable=True):
        pytest.skip("oracle/_ref not built")
    s1 = synth.make_reads(40, seed=90, read_len=30, dup_frac=0.4)
    fq, fa = synth.to_fastq(s1), synth.to_fasta(s1)
    cases = {
        "crlf_fastq": (fq.replace(b"\n", b"\r\n"), "fastq"),
        "crlf_fasta": (fa.replace(b"\n", b"\r\n"), "fasta"),
        "multiline_fasta": (b"".join(b">r%d\n" % i + s[:15] + b"\n" + s[15:] + b"\n" for i, s in enumerate(s1)), "fasta"),
        "fastq_as_fasta": (fq, "fasta"),
        "fasta_as_fastq": (fa, "fastq"),
        "tabs_in_seq": (fq.replace(b"A", b"\t", 3), "fastq"),
        "empty_seq_lines": (b"@a\n\n+\n\n@b\n\n+\n\n@c\nA\n+\nI\n", "fastq"),
        "plus_with_id": (b"".join(b"@r%d\n" % i + s + b"\n+r%d\n" % i + b"I" * len(s) + b"\n" for i, s in enumerate(s1)), "fastq"),
        "binary_garbage": (bytes(range(256)) * 20, "fastq"),
        "only_newlines": (b"\n" * 50, "fastq"),
        "at_in_quality": (b"".join(b"@r%d\n" % i + s + b"\n+\n" + b"@" * len(s) + b"\n" for i, s in enumerate(s1)), "fastq"),
    }
    env = dict(os.environ, LD_LIBRARY_PATH=str(FAKE_DIR), FQD_IO_THREADS="4", FQD_BLOCK_BYTES="4096")
    for name, (data, fmt) in cases.items():
        for flags in (["--fast"], ["--compare-seq", "tight"], ["--compare-seq", "loose"], ["--compare-seq", "tail-hamming"])[: 2 if THIN else 4][(1 if THIN and name in ("multiline_fasta", "at_in_quality") else 0):]:
            d = tmp_path / name
            shutil.rmtree(d, ignore_errors=True)
            d.mkdir()
            (d / "a").write_bytes(data)
            common = ["-v", "--format", fmt, *flags]
            ref_bin = oracle.REF_STABLE_BIN if flags[0] == "--compare-seq" else oracle.REF_BIN
            ref = subprocess.run([str(ref_bin), "-i", "a", "-o", "r1", *common], capture_output=True, cwd=d)
            ours = subprocess.run([str(EXE), "-i", "a", "-o", "o1", *common], capture_output=True, cwd=d, env=env)
            ours_err = b"".join(l for l in ours.stderr.splitlines(keepends=True) if not l.startswith(b"[fake_fqd]"))
            where = (name, flags)
            assert (ours.returncode, ours.stdout, ours_err) == (ref.returncode, ref.stdout, ref.stderr), where
            assert (d / "o1").exists() == (d / "r1").exists(), where
            if (d / "o1").exists():
                assert (d / "o1").read_bytes() == (d / "r1").read_bytes(), where


def test_fifo_input_and_fifo_output_in_a_sequence_mode(tmp_path, oracle):
    """Pipes on both sides of a whole-input mode: the input has no size (tables start small and grow), the output has no
    offsets (sequential writes)."""
    import threading
    seqs = synth.make_reads(100000, seed=71, read_len=40, dup_frac=0.4, var_len=True)
    buf = synth.to_fastq(seqs)
    exp, _, est = oracle.run_oracle("tight", oracle.FASTQ, buf)
    os.mkfifo(tmp_path / "in.fifo")
    os.mkfifo(tmp_path / "out.fifo")
    got = {}
    def feed():
        with open(tmp_path / "in.fifo", "wb") as f:
            f.write(buf)
    def drain():
        with open(tmp_path / "out.fifo", "rb") as f:
            got["out"] = f.read()
    th = [threading.Thread(target=feed), threading.Thread(target=drain)]
    for t in th:
        t.start()
    p = run("-i", tmp_path / "in.fifo", "-o", tmp_path / "out.fifo", "--compare-seq", "tight", "-v")
    for t in th:
        t.join(timeout=60)
    assert p.returncode == 0, p.stderr
    assert got["out"] == exp
    assert p.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"

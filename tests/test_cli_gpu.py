"""The drop-in binary, driven exactly like the reference's own test-suite drives the reference
(test/test_basic.py, test_fast.py, test_seq.py, test_unordered.py: subprocess + byte comparison of the outputs)."""
import gzip
import subprocess
from pathlib import Path

import pytest

import synth

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "fastq-dupaway_b200" / "host" / "fastq-dupaway"
FIX = ROOT / "tests" / "golden" / "ref_fixtures"


def run(*args, cwd=None, env=None):
    import os
    return subprocess.run([str(EXE), *map(str, args)], capture_output=True, text=True, cwd=cwd,
                          env=dict(os.environ, **env) if env else None)


def test_exe_available():
    assert EXE.exists(), "build the host binary: make -C fastq-dupaway_b200/host"


def test_help():
    # test/test_basic.py:13-22 - help goes to stderr, exit status 1
    res = run("-h")
    assert res.returncode == 1
    assert res.stderr.startswith("fastq-dupaway V")


@pytest.mark.parametrize("argv,msg", [
    (["-i", "a"], "the option '--output-1' is required but missing"),
    (["-i", "a", "-o", "b", "-u", "c"], "Both input-2 and output-2 arguments are required for paired-end mode!"),
    (["-i", "a", "-o", "b", "-u", "a", "-p", "c"], "Paired input files should not be the same file!"),
    (["-i", "a", "-o", "b", "--format", "bam"], 'Only "fastq" or "fasta" file formats are supported!'),
    (["-i", "a", "-o", "b", "--compare-seq", "fuzzy"], "Unsupported compare-seq type provided!"),
    (["-i", "a", "-o", "b", "-m", "100"], "Value of unsupported range provided for --mem-limit option!"),
    (["-i", "a", "-o", "b", "--fast", "--distance", "1"], "--fast mode was enabled, but argument(s) for sequence-based mode were provided!"),
    (["-i", "a", "-o", "b", "--unordered"], "--unordered argument can only be used with --fast mode!"),
    (["-i", "a", "-o", "b", "--fast", "--unordered"], "--unordered argument can only be used with paired inputs!"),
])
def test_argument_rules(argv, msg):
    # src/main.cpp:93-172
    res = run(*argv)
    assert res.returncode == 1
    assert res.stderr == "An error occured during arguments parsing:\n" + msg + "\n"


pytestmark_gpu = pytest.mark.gpu


@pytest.mark.gpu
def test_single_fast(tmp_path):
    out = tmp_path / "single_fast.fa"
    res = run("-i", FIX / "inputs" / "single_fast.fa", "-o", out, "--format", "fasta", "--fast")
    assert res.returncode == 0, res.stderr
    assert out.read_bytes() == (FIX / "expected" / "single_fast.fa").read_bytes()


@pytest.mark.gpu
def test_paired_fast(tmp_path):
    o1, o2 = tmp_path / "r1.fa", tmp_path / "r2.fa"
    res = run("-i", FIX / "inputs" / "paired_fast_r1.fa", "-u", FIX / "inputs" / "paired_fast_r2.fa", "-o", o1, "-p", o2,
              "--format", "fasta", "--fast", "-v")
    assert res.returncode == 0, res.stderr
    assert o1.read_bytes() == (FIX / "expected" / "paired_fast_r1.fa").read_bytes()
    assert o2.read_bytes() == (FIX / "expected" / "paired_fast_r2.fa").read_bytes()
    assert res.stdout == "10 read pairs processed, out of which 3 duplicates were removed.\n"


@pytest.mark.gpu
@pytest.mark.parametrize("filename,cli_args", [
    ("single_tight.fa", ["--format", "fasta"]),
    ("single_loose.fa", ["--format", "fasta", "--compare-seq", "loose"]),
    ("single_hamming.fa", ["--format", "fasta", "--compare-seq", "tail-hamming", "--distance", "1"]),
])
def test_single_fasta_seq_modes(tmp_path, filename, cli_args):
    out = tmp_path / filename
    res = run("-i", FIX / "inputs" / filename, "-o", out, *cli_args)
    assert res.returncode == 0, res.stderr
    assert out.read_bytes() == (FIX / "expected" / filename).read_bytes()


@pytest.mark.gpu
def test_paired_fasta_tight(tmp_path):
    o1, o2 = tmp_path / "r1.fa", tmp_path / "r2.fa"
    res = run("-i", FIX / "inputs" / "paired_tight_r1.fa", "-u", FIX / "inputs" / "paired_tight_r2.fa", "-o", o1, "-p", o2, "--format", "fasta")
    assert res.returncode == 0, res.stderr
    assert o1.read_bytes() == (FIX / "expected" / "paired_tight_r1.fa").read_bytes()
    assert o2.read_bytes() == (FIX / "expected" / "paired_tight_r2.fa").read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["shuffled", "skewed", "deletion", "interleaved", "not_overlapped"])
def test_unordered(tmp_path, name):
    o1, o2 = tmp_path / "r1.fa", tmp_path / "r2.fa"
    res = run("-i", FIX / "inputs" / f"unordered_{name}_r1.fa", "-u", FIX / "inputs" / f"unordered_{name}_r2.fa", "-o", o1, "-p", o2,
              "--format", "fasta", "--fast", "--unordered")
    assert res.returncode == 0, res.stderr
    assert o1.read_bytes() == (FIX / "expected" / f"unordered_{name}_r1.fa").read_bytes()
    assert o2.read_bytes() == (FIX / "expected" / f"unordered_{name}_r2.fa").read_bytes()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fast", "tight", "loose", "tail-hamming"])
@pytest.mark.parametrize("gz", [False, True])
def test_fastq_against_oracle_through_files(tmp_path, oracle, mode, gz):
    """FASTQ, plain and .gz (by extension, independently for input and output), SE, with -v; multi-block input."""
    seqs = synth.make_reads(60000, seed=41, read_len=100, var_len=True, n_frac=0.02, prefix_frac=0.2, sub_frac=0.2, dup_frac=0.4)
    buf = synth.to_fastq(seqs)
    ext = ".fq.gz" if gz else ".fq"
    inp, out = tmp_path / ("in" + ext), tmp_path / ("out" + ext)
    inp.write_bytes(gzip.compress(buf, 1) if gz else buf)
    args = ["-i", inp, "-o", out, "-v", "-m", "500"]
    args += ["--fast"] if mode == "fast" else ["--compare-seq", mode]
    res = run(*args)
    assert res.returncode == 0, res.stderr
    exp, _, est = oracle.run_oracle(mode, oracle.FASTQ, buf)
    got = gzip.decompress(out.read_bytes()) if gz else out.read_bytes()
    assert got == exp
    assert res.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"


@pytest.mark.gpu
def test_paired_fastq_fast_and_unordered_through_files(tmp_path, oracle):
    import numpy as np
    s1, s2 = synth.make_pair(20000, seed=42, read_len=80)
    b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    (tmp_path / "a.fq").write_bytes(b1)
    (tmp_path / "b.fq").write_bytes(b2)
    res = run("-i", tmp_path / "a.fq", "-u", tmp_path / "b.fq", "-o", tmp_path / "o1.fq", "-p", tmp_path / "o2.fq", "--fast", "-v")
    assert res.returncode == 0, res.stderr
    e1, e2, est = oracle.run_oracle("fast", oracle.FASTQ, b1, b2)
    assert (tmp_path / "o1.fq").read_bytes() == e1 and (tmp_path / "o2.fq").read_bytes() == e2
    assert res.stdout == f"{est.total} read pairs processed, out of which {est.dups} duplicates were removed.\n"
    # shuffle R2 and drop a few records: --unordered
    rng = np.random.default_rng(43)
    recs = [b2[i:i + 0] for i in range(0)]
    lines = b2.split(b"\n")[:-1]
    recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines), 4)]
    recs = [r for r in recs if rng.random() > 0.05]
    rng.shuffle(recs)
    b2s = b"".join(recs)
    (tmp_path / "bs.fq").write_bytes(b2s)
    res = run("-i", tmp_path / "a.fq", "-u", tmp_path / "bs.fq", "-o", tmp_path / "u1.fq", "-p", tmp_path / "u2.fq", "--fast", "--unordered", "-v")
    assert res.returncode == 0, res.stderr
    e1, e2, est = oracle.run_oracle("fast", oracle.FASTQ, b1, b2s, unordered=True)
    assert (tmp_path / "u1.fq").read_bytes() == e1 and (tmp_path / "u2.fq").read_bytes() == e2
    assert res.stdout == (f"{est.total} valid read pairs processed, out of which {est.dups} duplicates were removed.\n"
                          f"{est.unmatched} Non-matching entries from both files were skipped.\n")


@pytest.mark.gpu
def test_error_messages_and_exit_codes(tmp_path):
    def go(content, *extra):
        p = tmp_path / "e.fq"
        p.write_bytes(content)
        return run("-i", p, "-o", tmp_path / "e.out", "--fast", *extra)
    banner = "An error occured during fastq-dupaway execution:\n"
    r = go(b"")
    assert r.returncode == 1 and r.stderr == banner + "Not enough memory to read a single object!\n"
    r = go(b"@a\nACGT\n+\nFFFF\n@b\nACXT\n+\nFFFF\n")
    assert r.returncode == 1
    assert r.stderr == "Error: unknown character in DNA sequence: X\n" + banner + "Supported sequence character set: {A, N, C, G, T}!\n"
    assert (tmp_path / "e.out").read_bytes() == b"@a\nACGT\n+\nFFFF\n"
    r = go(b"@a\nACXT\n+\nFFFF\n@b\nAAAA\n+\nFFFF\n")      # the reference writes record 0 before keying it
    assert r.returncode == 1 and (tmp_path / "e.out").read_bytes() == b"@a\nACXT\n+\nFFFF\n"
    r = go(b"xa\nACGT\n+\nFFFF\n")
    assert r.returncode == 1 and r.stderr == "Invalid record start character: x\n" + banner + "Fastq record should start with @ symbol!\n"
    r = go(b"@a\nACGT\n+\nFFF\n")
    assert r.returncode == 1
    assert r.stderr == ("Found sequence ACGT of length 5 and quality string FFF of length 4\n" + banner +
                        "Sequence and Quality fields of Fastq record should have the same length!\n")
    r = run("-i", tmp_path / "missing.fq", "-o", tmp_path / "x.out", "--fast")
    assert r.returncode == 1 and r.stderr == f"Cannot open file {tmp_path / 'missing.fq'}\n" + banner + "File does not exist or cannot be opened!\n"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["tight", "loose"])
def test_write_clusters_through_files(tmp_path, oracle, mode):
    """--write-clusters: <out>.clusters next to every output file (src/file_utils.cpp:98-112)."""
    s1, s2 = synth.make_pair(5000, seed=81, read_len=50, var_len=True, prefix_frac=0.3, dup_frac=0.5)
    b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    (tmp_path / "r1.fq").write_bytes(b1)
    (tmp_path / "r2.fq").write_bytes(b2)
    res = run("-i", tmp_path / "r1.fq", "-u", tmp_path / "r2.fq", "-o", tmp_path / "o1.fq", "-p", tmp_path / "o2.fq",
              "--compare-seq", mode, "--write-clusters")
    assert res.returncode == 0, res.stderr
    texts, _ = oracle.cluster_text(mode, oracle.FASTQ, b1, b2)
    e1, e2, _ = oracle.run_oracle(mode, oracle.FASTQ, b1, b2)
    assert (tmp_path / "o1.fq").read_bytes() == e1 and (tmp_path / "o2.fq").read_bytes() == e2
    assert (tmp_path / "o1.fq.clusters").read_bytes() == texts[0]
    assert (tmp_path / "o2.fq.clusters").read_bytes() == texts[1]


@pytest.mark.gpu
def test_arbitrary_bytes_through_files(tmp_path, oracle):
    """Lower case / IUPAC symbols in sequence-based mode: accepted and ordered like the reference does (any byte)."""
    seqs = synth.make_reads(20000, seed=95, read_len=80, var_len=True, prefix_frac=0.2, sub_frac=0.2, dup_frac=0.4, alphabet=b"ACGTNacgtnRYKM")
    buf = synth.to_fastq(seqs)
    (tmp_path / "in.fq").write_bytes(buf)
    res = run("-i", tmp_path / "in.fq", "-o", tmp_path / "out.fq", "--compare-seq", "tail-hamming", "-v")
    assert res.returncode == 0, res.stderr
    exp, _, est = oracle.run_oracle("tail-hamming", oracle.FASTQ, buf)
    assert (tmp_path / "out.fq").read_bytes() == exp
    assert res.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fast", "tight"])
def test_paired_multi_member_gzip_and_bgzf_inputs(tmp_path, oracle, mode):
    """R1 as a multi-member archive, R2 as BGZF (both inflated member-parallel on the host, pargz.hpp), .gz outputs
    (deflated in parallel, multi-member), many small blocks (FQD_BLOCK_BYTES) so that blocks are recycled behind the
    asynchronous writers; byte parity of the decompressed outputs with the oracle."""
    from test_host_io import bgzf, members
    s1, s2 = synth.make_pair(120000, seed=77, read_len=100)
    b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    (tmp_path / "a.fq.gz").write_bytes(members(b1, [1_500_000, 20_000]))
    (tmp_path / "b.fq.gz").write_bytes(bgzf(b2))
    args = ["-i", tmp_path / "a.fq.gz", "-u", tmp_path / "b.fq.gz", "-o", tmp_path / "o1.fq.gz", "-p", tmp_path / "o2.fq.gz", "-v"]
    args += ["--fast"] if mode == "fast" else ["--compare-seq", mode]
    res = run(*args, env={"FQD_BLOCK_BYTES": str(1 << 20)})
    assert res.returncode == 0, res.stderr
    e1, e2, est = oracle.run_oracle(mode, oracle.FASTQ, b1, b2)
    assert gzip.decompress((tmp_path / "o1.fq.gz").read_bytes()) == e1
    assert gzip.decompress((tmp_path / "o2.fq.gz").read_bytes()) == e2
    assert res.stdout == f"{est.total} read pairs processed, out of which {est.dups} duplicates were removed.\n"


@pytest.mark.gpu
def test_single_member_gzip_inflated_block_parallel(tmp_path, oracle):
    """`gzip reads.fq` writes ONE member; the host cuts its deflate stream into chunks and inflates them on all
    threads (pinflate.hpp).  Small chunk sizes so that this small archive is cut into dozens of pieces."""
    from test_host_io import SMALL, deflate_gz
    seqs = synth.make_reads(80000, seed=78, read_len=100, var_len=True, n_frac=0.02, dup_frac=0.4)
    buf = synth.to_fastq(seqs)
    (tmp_path / "in.fq.gz").write_bytes(deflate_gz(buf, 6))
    res = run("-i", tmp_path / "in.fq.gz", "-o", tmp_path / "out.fq", "--fast", "-v", env=dict(SMALL, FQD_BLOCK_BYTES=str(1 << 20)))
    assert res.returncode == 0, res.stderr
    exp, _, est = oracle.run_oracle("fast", oracle.FASTQ, buf)
    assert (tmp_path / "out.fq").read_bytes() == exp
    assert res.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"


@pytest.mark.gpu
@pytest.mark.parametrize("devices,block", [("0,0", 1 << 16), ("0,0,0", 1 << 18), ("0,0", None)])
@pytest.mark.parametrize("paired", [False, True])
def test_fast_sharded_over_several_engines(tmp_path, oracle, devices, block, paired):
    """FQD_DEVICES: the duplicate set of `--fast` is sharded by hash range over several engines of ONE process (here: on the
    same GPU - the protocol is the same, rows travel through the copy engines into the owner's key store); chunks are dealt
    round-robin, survivors are written in input order: outputs and -v line identical to the oracle's."""
    s1, s2 = synth.make_pair(60000, seed=95, read_len=100, var_len=True, n_frac=0.02, dup_frac=0.4)
    s2 = s2[:-7]
    b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    (tmp_path / "a.fq").write_bytes(b1)
    (tmp_path / "b.fq").write_bytes(b2)
    env = {"FQD_DEVICES": devices}
    if block:
        env["FQD_BLOCK_BYTES"] = str(block)
    if paired:
        res = run("-i", tmp_path / "a.fq", "-u", tmp_path / "b.fq", "-o", tmp_path / "o1.fq", "-p", tmp_path / "o2.fq", "--fast", "-v", env=env)
        e1, e2, est = oracle.run_oracle("fast", oracle.FASTQ, b1, b2)
        assert res.returncode == 0, res.stderr
        assert (tmp_path / "o1.fq").read_bytes() == e1 and (tmp_path / "o2.fq").read_bytes() == e2
        assert res.stdout == f"{est.total} read pairs processed, out of which {est.dups} duplicates were removed.\n"
    else:
        res = run("-i", tmp_path / "a.fq", "-o", tmp_path / "o1.fq", "--fast", "-v", env=env)
        e1, _, est = oracle.run_oracle("fast", oracle.FASTQ, b1)
        assert res.returncode == 0, res.stderr
        assert (tmp_path / "o1.fq").read_bytes() == e1
        assert res.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["base", "start", "length"])
def test_fast_sharded_error_paths_match_the_single_engine_binary(tmp_path, kind):
    """A malformed record in the middle of the input: the sharded driver stops where the single-engine driver stops (same exit
    status, stderr, output bytes) - both follow the reference's order of events (tests/test_differential_gpu.py pins the latter)."""
    seqs = synth.make_reads(9000, seed=96, read_len=80, dup_frac=0.3)
    recs = [synth.to_fastq([s], ids=[b"@r.%d" % i]) for i, s in enumerate(seqs)]
    bad = 5432
    r = recs[bad].split(b"\n")
    if kind == "base":
        r[1] = r[1][:10] + b"Z" + r[1][11:]
    elif kind == "start":
        r[0] = b"x" + r[0][1:]
    else:
        r[3] = r[3][:-3]
    recs[bad] = b"\n".join(r)
    (tmp_path / "a.fq").write_bytes(b"".join(recs))
    outs = []
    for tag, env in (("one", {"FQD_BLOCK_BYTES": str(1 << 16)}), ("many", {"FQD_BLOCK_BYTES": str(1 << 16), "FQD_DEVICES": "0,0,0"})):
        res = run("-i", tmp_path / "a.fq", "-o", tmp_path / f"{tag}.fq", "--fast", "-v", env=env)
        outs.append((res.returncode, res.stdout, res.stderr, (tmp_path / f"{tag}.fq").read_bytes()))
    assert outs[0][0] == 1
    assert outs[0] == outs[1]

"""The drop-in binary's HOST logic on the CPU: the real `fastq-dupaway` executable, with a test double of the C ABI
(tests/fake_engine/fake_fqd.cpp: newline counting + a std::unordered_set) preloaded in place of libfqd_cuda.so.

What runs here is everything AROUND the engine - options, plain / gzip readers, rings of pinned blocks, tail carry,
paired lock-step, restarts after wrong capacity / row-width estimates, asynchronous writers, gzip outputs, -v lines -
checked byte for byte against the oracle.  The engine itself (and the error-order rules it implements) is tested on the
GPU (tests/test_cli_gpu.py); the product has no CPU path: tests/test_abi.py checks that the binary fails without a GPU.
"""
import gzip
import os
import subprocess
import sys
from pathlib import Path

import pytest

import synth
from test_host_io import SMALL, bgzf, deflate_gz, members

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "fastq-dupaway_b200" / "host" / "fastq-dupaway"
sys.path.insert(0, str(ROOT / "tests" / "fake_engine"))
from build import BUILD as FAKE_DIR, build_fake  # noqa: E402


# Every position by default.  tests/test_differential_gpu.py thins the sweeps (each run of the real binary pays ~1 s of
# CUDA start-up): first, last and every SWEEP_STRIDE-th position; FQD_DIFF_FULL=1 restores every position there.
SWEEP_STRIDE = 1


def _sweep(n):
    if SWEEP_STRIDE > 1:            # thinned (GPU box): one position early, one in the last block
        return [1, n - 2]
    return list(range(n))


@pytest.fixture(scope="module", autouse=True)

def fake_engine():
    build_fake()
    # the host binary needs the real library only to LINK; build it if the tree is fresh
    if not EXE.exists():
        subprocess.run(["make", "-s", "-C", str(EXE.parent)], check=True)


def run(*args, env=None):
    e = dict(os.environ, LD_LIBRARY_PATH=str(FAKE_DIR), FQD_IO_THREADS="4")
    e.update(env or {})
    p = subprocess.run([str(EXE), *map(str, args)], capture_output=True, text=True, env=e, timeout=300)
    assert "[fake_fqd] TEST DOUBLE" in p.stderr or p.returncode != 0 or "-h" in args
    p.stderr = "".join(l + "\n" for l in p.stderr.splitlines() if not l.startswith("[fake_fqd]"))
    return p


@pytest.mark.parametrize("threads", ["1", "4"])
@pytest.mark.parametrize("block", [4096, 1 << 16, None])
def test_single_end_blocks_and_tail_carry(tmp_path, oracle, threads, block):
    seqs = synth.make_reads(30000, seed=5, read_len=100, var_len=True, n_frac=0.02, dup_frac=0.4)
    buf = synth.to_fastq(seqs)
    (tmp_path / "in.fq").write_bytes(buf)
    env = {"FQD_IO_THREADS": threads}
    if block:
        env["FQD_BLOCK_BYTES"] = str(block)
    res = run("-i", tmp_path / "in.fq", "-o", tmp_path / "out.fq", "--fast", "-v", env=env)
    assert res.returncode == 0, res.stderr
    exp, _, est = oracle.run_oracle("fast", oracle.FASTQ, buf)
    assert (tmp_path / "out.fq").read_bytes() == exp
    assert res.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"


def test_paired_gzip_inputs_and_outputs_files_of_different_length(tmp_path, oracle):
    s1, s2 = synth.make_pair(40000, seed=6, read_len=80)
    b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2[:35000], mate=2)       # R2 is shorter: stop there
    (tmp_path / "a.fq.gz").write_bytes(members(b1, [400_000, 30_000]))
    (tmp_path / "b.fq.gz").write_bytes(deflate_gz(b2, 6))                         # one member: block-parallel inflate
    res = run("-i", tmp_path / "a.fq.gz", "-u", tmp_path / "b.fq.gz", "-o", tmp_path / "o1.fq.gz", "-p", tmp_path / "o2.fq.gz",
              "--fast", "-v", env=dict(SMALL, FQD_BLOCK_BYTES=str(1 << 18)))
    assert res.returncode == 0, res.stderr
    e1, e2, est = oracle.run_oracle("fast", oracle.FASTQ, b1, b2)
    assert gzip.decompress((tmp_path / "o1.fq.gz").read_bytes()) == e1
    assert gzip.decompress((tmp_path / "o2.fq.gz").read_bytes()) == e2
    assert res.stdout == f"{est.total} read pairs processed, out of which {est.dups} duplicates were removed.\n"


def test_fasta_and_bgzf(tmp_path, oracle):
    seqs = synth.make_reads(20000, seed=7, read_len=60, var_len=True, dup_frac=0.3)
    buf = synth.to_fasta(seqs)
    (tmp_path / "in.fa.gz").write_bytes(bgzf(buf))
    res = run("-i", tmp_path / "in.fa.gz", "-o", tmp_path / "out.fa", "--format", "fasta", "--fast", "-v", env={"FQD_BLOCK_BYTES": str(1 << 16)})
    assert res.returncode == 0, res.stderr
    exp, _, est = oracle.run_oracle("fast", oracle.FASTA, buf)
    assert (tmp_path / "out.fa").read_bytes() == exp


def test_restart_when_the_capacity_estimate_was_too_small(tmp_path, oracle):
    """The double pretends its key store is 4 x smaller than asked for: the run hits FQD_ERR_CAPACITY after survivors
    have been written, starts over with doubled tables (twice), and the final output is still exact."""
    seqs = synth.make_reads(150000, seed=8, read_len=50, dup_frac=0.2)
    buf = synth.to_fastq(seqs)
    (tmp_path / "in.fq").write_bytes(buf)
    res = run("-i", tmp_path / "in.fq", "-o", tmp_path / "out.fq", "--fast", "-v",
              env={"FAKE_FQD_SHRINK": "4", "FQD_BLOCK_BYTES": str(1 << 20), "FQD_TRACE": "1"})
    assert res.returncode == 0, res.stderr
    assert res.stderr.count("engine created") == 3            # two restarts
    exp, _, est = oracle.run_oracle("fast", oracle.FASTQ, buf)
    assert (tmp_path / "out.fq").read_bytes() == exp
    assert res.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"


def test_restart_when_a_later_read_is_longer_than_the_sampled_ones(tmp_path, oracle):
    short = synth.make_reads(3000, seed=9, read_len=40, dup_frac=0.2)
    long_ = synth.make_reads(3000, seed=10, read_len=130, dup_frac=0.2)
    buf = synth.to_fastq(short + long_)
    (tmp_path / "in.fq").write_bytes(buf)
    res = run("-i", tmp_path / "in.fq", "-o", tmp_path / "out.fq", "--fast", "-v", env={"FQD_BLOCK_BYTES": str(1 << 14), "FQD_TRACE": "1"})
    assert res.returncode == 0, res.stderr
    assert res.stderr.count("engine created") >= 2            # started over with wider key rows
    exp, _, est = oracle.run_oracle("fast", oracle.FASTQ, buf)
    assert (tmp_path / "out.fq").read_bytes() == exp


def test_errors_reach_the_user_with_the_reference_wording(tmp_path):
    # empty input (src/bufferedinput.hpp:82-85)
    (tmp_path / "e.fq").write_bytes(b"")
    res = run("-i", tmp_path / "e.fq", "-o", tmp_path / "o.fq", "--fast")
    assert res.returncode == 1 and res.stderr == "An error occured during fastq-dupaway execution:\nNot enough memory to read a single object!\n"
    # missing input (src/file_utils.hpp:110-121); the output file has been created already, like the reference
    res = run("-i", tmp_path / "nope.fq", "-o", tmp_path / "o2.fq", "--fast")
    assert res.returncode == 1 and "Cannot open file" in res.stderr and (tmp_path / "o2.fq").exists()
    # wrong first byte (src/fastqview.cpp:121-126)
    (tmp_path / "x.fq").write_bytes(b">r\nACGT\n")
    res = run("-i", tmp_path / "x.fq", "-o", tmp_path / "o3.fq", "--fast")
    assert res.returncode == 1 and "Invalid record start character: >" in res.stderr
    assert "Fastq record should start with @ symbol!" in res.stderr
    # a base outside {A,C,G,T,N} in a later block (src/seq_utils.cpp:17-19): what came before is written
    good = synth.to_fastq(synth.make_reads(2000, seed=11, read_len=50, dup_frac=0.0))
    (tmp_path / "b.fq").write_bytes(good + b"@bad\nACGU\n+\nIIII\n" + good)
    res = run("-i", tmp_path / "b.fq", "-o", tmp_path / "o4.fq", "--fast", env={"FQD_BLOCK_BYTES": str(1 << 14)})
    assert res.returncode == 1
    assert "Error: unknown character in DNA sequence: U" in res.stderr
    assert "Supported sequence character set: {A, N, C, G, T}!" in res.stderr
    assert (tmp_path / "o4.fq").read_bytes() == good
    # corrupt gzip input
    blob = bytearray(deflate_gz(good, 6))
    blob[len(blob) // 2] ^= 0x20
    (tmp_path / "c.fq.gz").write_bytes(bytes(blob))
    res = run("-i", tmp_path / "c.fq.gz", "-o", tmp_path / "o5.fq", "--fast")
    assert res.returncode == 1 and "gzip error" in res.stderr


def _records(n, seed, mate=1, read_len=100):
    seqs = synth.make_reads(n, seed=seed, read_len=read_len, dup_frac=0.3)
    buf = synth.to_fastq(seqs, mate=mate)
    lines = buf.split(b"\n")[:-1]
    return [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines), 4)]


def _damage(rec, kind):
    l = rec.split(b"\n")
    if kind == "start":
        l[0] = b"X" + l[0][1:]
    elif kind == "length":
        l[3] = l[3][:-3]
    elif kind == "base":
        l[1] = l[1][:10] + b"U" + l[1][11:]
    return b"\n".join(l)


@pytest.mark.parametrize("kind", ["start", "length", "base"])
def test_malformed_record_at_every_position_matches_the_reference_binary(tmp_path, oracle, kind):
    """One malformed record, at every index of a 40-record file that spans ten 1 KiB..4 KiB blocks here and one block in
    the reference: exit status, stderr and the bytes written before the stop must be the reference binary's.  This is
    the lazy pre-parse rule (src/bufferedinput.hpp:90-103) carried across block boundaries by dup_remover.cpp - the tail
    check, the peeked next block, the record dropped in front of a bad start."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/fastq-dupaway not built")
    recs = _records(40, seed=30)
    for pos in _sweep(40):
        data = b"".join(recs[:pos] + [_damage(recs[pos], kind)] + recs[pos + 1:])
        inp = tmp_path / "in.fq"
        inp.write_bytes(data)
        ref = subprocess.run([str(oracle.REF_BIN), "-i", str(inp), "-o", str(tmp_path / "ref.fq"), "--fast", "-v"],
                             capture_output=True, text=True, cwd=tmp_path)
        for block in ((4096, 1 << 20) if SWEEP_STRIDE == 1 else (4096,)):
            ours = run("-i", inp, "-o", tmp_path / "ours.fq", "--fast", "-v", env={"FQD_BLOCK_BYTES": str(block)})
            assert ours.returncode == ref.returncode == 1, (pos, block)
            assert ours.stderr == ref.stderr, (pos, block)
            assert ours.stdout == ref.stdout
            assert (tmp_path / "ours.fq").read_bytes() == (tmp_path / "ref.fq").read_bytes(), (pos, block)


@pytest.mark.parametrize("kind", ["start", "length", "base"])
@pytest.mark.parametrize("bad_mate", [0, 1])
def test_malformed_record_in_paired_input_matches_the_reference_binary(tmp_path, oracle, kind, bad_mate):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/fastq-dupaway not built")
    r = [_records(24, seed=31, mate=1), _records(24, seed=32, mate=2, read_len=80)]
    for pos in _sweep(24):
        files = []
        for m in (0, 1):
            recs = list(r[m])
            if m == bad_mate:
                recs[pos] = _damage(recs[pos], kind)
            f = tmp_path / f"in{m}.fq"
            f.write_bytes(b"".join(recs))
            files.append(f)
        ref = subprocess.run([str(oracle.REF_BIN), "-i", str(files[0]), "-u", str(files[1]), "-o", str(tmp_path / "r1.fq"),
                              "-p", str(tmp_path / "r2.fq"), "--fast", "-v"], capture_output=True, text=True, cwd=tmp_path)
        ours = run("-i", files[0], "-u", files[1], "-o", tmp_path / "o1.fq", "-p", tmp_path / "o2.fq", "--fast", "-v",
                   env={"FQD_BLOCK_BYTES": "4096"})
        assert ours.returncode == ref.returncode == 1, pos
        assert ours.stderr == ref.stderr, pos
        assert (tmp_path / "o1.fq").read_bytes() == (tmp_path / "r1.fq").read_bytes(), pos
        assert (tmp_path / "o2.fq").read_bytes() == (tmp_path / "r2.fq").read_bytes(), pos


def test_input_cut_anywhere_in_the_last_record_matches_the_reference_binary(tmp_path, oracle):
    """The file ends at every byte of its last record (and a little before it), single-end and in either mate of a
    pair: the binary, the oracle and the reference binary agree on exit status, stderr, -v line and output bytes."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/fastq-dupaway not built")
    r1, r2 = _records(12, seed=33), _records(12, seed=34, mate=2, read_len=70)
    f1, f2 = b"".join(r1), b"".join(r2)
    for cut in range(len(f1) - len(r1[-1]) - 2, len(f1) + 1, 3 if SWEEP_STRIDE == 1 else 61):
        for paired in (False, True):
            b1 = f1[:cut]
            (tmp_path / "a.fq").write_bytes(b1)
            (tmp_path / "b.fq").write_bytes(f2)
            io_ref = ["-i", "a.fq", "-o", "r1.fq"] + (["-u", "b.fq", "-p", "r2.fq"] if paired else [])
            io_our = ["-i", tmp_path / "a.fq", "-o", tmp_path / "o1.fq"] + (["-u", tmp_path / "b.fq", "-p", tmp_path / "o2.fq"] if paired else [])
            ref = subprocess.run([str(oracle.REF_BIN), *io_ref, "--fast", "-v"], capture_output=True, text=True, cwd=tmp_path)
            ours = run(*io_our, "--fast", "-v", env={"FQD_BLOCK_BYTES": "4096"})
            assert (ours.returncode, ours.stdout, ours.stderr) == (ref.returncode, ref.stdout, ref.stderr), (cut, paired)
            assert (tmp_path / "o1.fq").read_bytes() == (tmp_path / "r1.fq").read_bytes(), (cut, paired)
            e1, e2, st = oracle.run_oracle("fast", oracle.FASTQ, b1, f2 if paired else None)
            assert e1 == (tmp_path / "r1.fq").read_bytes() and (st.err != 0) == (ref.returncode != 0)
            if paired:
                assert (tmp_path / "o2.fq").read_bytes() == (tmp_path / "r2.fq").read_bytes() == e2, cut


@pytest.mark.parametrize("kind", ["start", "base", "empty", "lower"])
def test_malformed_fasta_record_at_every_position_matches_the_reference_binary(tmp_path, oracle, kind):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/fastq-dupaway not built")
    buf = synth.to_fasta(synth.make_reads(30, seed=40, read_len=90, dup_frac=0.3))
    lines = buf.split(b"\n")[:-1]
    recs = [b"\n".join(lines[i:i + 2]) + b"\n" for i in range(0, len(lines), 2)]

    def damage(rec):
        l = rec.split(b"\n")
        if kind == "start":
            l[0] = b"@" + l[0][1:]
        elif kind == "base":
            l[1] = l[1][:10] + b"U" + l[1][11:]
        elif kind == "empty":
            l[1] = b""
        else:
            l[1] = l[1].lower()
        return b"\n".join(l)
    for pos in _sweep(30):
        data = b"".join(recs[:pos] + [damage(recs[pos])] + recs[pos + 1:])
        (tmp_path / "in.fa").write_bytes(data)
        ref = subprocess.run([str(oracle.REF_BIN), "-i", "in.fa", "-o", "ref.fa", "--fast", "-v", "--format", "fasta"],
                             capture_output=True, text=True, cwd=tmp_path)
        ours = run("-i", tmp_path / "in.fa", "-o", tmp_path / "ours.fa", "--fast", "-v", "--format", "fasta", env={"FQD_BLOCK_BYTES": "4096"})
        assert (ours.returncode, ours.stdout, ours.stderr) == (ref.returncode, ref.stdout, ref.stderr), pos
        assert (tmp_path / "ours.fa").read_bytes() == (tmp_path / "ref.fa").read_bytes(), pos
        e1, _, st = oracle.run_oracle("fast", oracle.FASTA, data)
        assert e1 == (tmp_path / "ref.fa").read_bytes(), pos


def test_two_malformed_records_report_the_one_the_reference_meets_first(tmp_path, oracle):
    """150 random jobs with TWO malformed records (any kinds, any positions, either mate, small or large blocks): which
    error is reported, and what has been written by then, follows the reference's order of events - the pre-parse of
    record k+1 of the left then the right mate comes before pair k is keyed (left mate first)."""
    import random
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/fastq-dupaway not built")
    rng = random.Random(5)
    r = [_records(24, seed=31, mate=1), _records(24, seed=32, mate=2, read_len=80)]
    for it in range(150 if SWEEP_STRIDE == 1 else 6):
        paired = rng.random() < 0.6
        recs = [list(r[0]), list(r[1])]
        for _ in range(2):
            m, pos = rng.randrange(2 if paired else 1), rng.randrange(24)
            recs[m][pos] = _damage(r[m][pos], rng.choice(["start", "length", "base"]))
        (tmp_path / "a.fq").write_bytes(b"".join(recs[0]))
        (tmp_path / "b.fq").write_bytes(b"".join(recs[1]))
        for f in ("r1.fq", "r2.fq", "o1.fq", "o2.fq"):
            (tmp_path / f).unlink(missing_ok=True)
        io_ref = ["-i", "a.fq", "-o", "r1.fq"] + (["-u", "b.fq", "-p", "r2.fq"] if paired else [])
        io_our = ["-i", tmp_path / "a.fq", "-o", tmp_path / "o1.fq"] + (["-u", tmp_path / "b.fq", "-p", tmp_path / "o2.fq"] if paired else [])
        ref = subprocess.run([str(oracle.REF_BIN), *io_ref, "--fast", "-v"], capture_output=True, text=True, cwd=tmp_path)
        ours = run(*io_our, "--fast", "-v", env={"FQD_BLOCK_BYTES": rng.choice(["4096", "1048576"])})
        assert (ours.returncode, ours.stderr) == (ref.returncode, ref.stderr), it
        assert (tmp_path / "o1.fq").read_bytes() == (tmp_path / "r1.fq").read_bytes(), it
        if paired:
            assert (tmp_path / "o2.fq").read_bytes() == (tmp_path / "r2.fq").read_bytes(), it


def _run_with_fifo_output(tmp_path, args, fifo_names):
    """Runs the binary with some outputs being FIFOs; returns (CompletedProcess, {name: bytes read from the FIFO})."""
    import threading
    got = {}
    def drain(name):
        with open(tmp_path / name, "rb") as f:
            got[name] = f.read()
    for n in fifo_names:
        os.mkfifo(tmp_path / n)
    th = [threading.Thread(target=drain, args=(n,)) for n in fifo_names]
    for t in th:
        t.start()
    p = run(*args)
    for t in th:
        t.join(timeout=60)
    return p, got


@pytest.mark.parametrize("paired", [False, True])
def test_non_seekable_outputs(tmp_path, oracle, paired):
    """-o /dev/stdout, -o <FIFO> (`-o >(pigz > x.gz)`): plain outputs are written with pwrite() at offsets on regular files
    only; anything else gets every byte through write() in order (round 1 wrote nothing and exited 0)."""
    s1, s2 = synth.make_pair(30000, seed=61, read_len=100, dup_frac=0.3)
    b1, b2 = synth.to_fastq(s1, mate=1), synth.to_fastq(s2, mate=2)
    (tmp_path / "a.fq").write_bytes(b1)
    (tmp_path / "b.fq").write_bytes(b2)
    if paired:
        e1, e2, _ = oracle.run_oracle("fast", oracle.FASTQ, b1, b2)
        p, got = _run_with_fifo_output(tmp_path, ["-i", tmp_path / "a.fq", "-u", tmp_path / "b.fq", "-o", tmp_path / "o1.fifo", "-p", tmp_path / "o2.fifo", "--fast"],
                                       ["o1.fifo", "o2.fifo"])
        assert p.returncode == 0, p.stderr
        assert got["o1.fifo"] == e1 and got["o2.fifo"] == e2
    else:
        e1, _, _ = oracle.run_oracle("fast", oracle.FASTQ, b1)
        p, got = _run_with_fifo_output(tmp_path, ["-i", tmp_path / "a.fq", "-o", tmp_path / "o.fifo", "--fast"], ["o.fifo"])
        assert p.returncode == 0, p.stderr
        assert got["o.fifo"] == e1
        # /dev/stdout of a process whose stdout is a pipe
        e = dict(os.environ, LD_LIBRARY_PATH=str(FAKE_DIR))
        q = subprocess.run([str(EXE), "-i", str(tmp_path / "a.fq"), "-o", "/dev/stdout", "--fast"], capture_output=True, env=e, timeout=120)
        assert q.returncode == 0 and q.stdout == e1


def test_failed_writes_fail_the_run(tmp_path):
    """A write error on a plain output (here: /dev/full) ends the run with the error banner and exit status 1."""
    (tmp_path / "a.fq").write_bytes(synth.to_fastq(synth.make_reads(2000, seed=62, read_len=80)))
    p = run("-i", tmp_path / "a.fq", "-o", "/dev/full", "--fast")
    assert p.returncode == 1 and "writing /dev/full failed" in p.stderr


def test_fifo_input_larger_than_the_first_table_size(tmp_path, oracle):
    """A FIFO has no size: the tables start at 65 536 records and must grow while the input streams in (round 1 tried to
    restart, which cannot work on a pipe: it hung / failed with 'Not enough memory to read a single object!')."""
    import threading
    seqs = synth.make_reads(120000, seed=63, read_len=60, dup_frac=0.3)
    buf = synth.to_fastq(seqs)
    exp, _, est = oracle.run_oracle("fast", oracle.FASTQ, buf)
    os.mkfifo(tmp_path / "in.fifo")
    def feed():
        with open(tmp_path / "in.fifo", "wb") as f:
            f.write(buf)
    t = threading.Thread(target=feed)
    t.start()
    p = run("-i", tmp_path / "in.fifo", "-o", tmp_path / "o.fq", "--fast", "-v")
    t.join(timeout=60)
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "o.fq").read_bytes() == exp
    assert p.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"

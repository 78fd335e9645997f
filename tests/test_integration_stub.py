"""INTEGRATION.md section 2, executed: a standalone C++ caller (tests/integration_stub/seq_stub.cpp - include/fqd.h and
nothing else of this repository) runs a sequence-based mode on the discarded-input path and writes the survivors and the
cluster file from its own mapping of the input.  On the CPU it links the test double of the ABI, under `-m gpu` the
product library; either way the bytes must be the oracle's."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

import synth

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "integration_stub" / "seq_stub.cpp"
sys.path.insert(0, str(ROOT / "tests" / "fake_engine"))
from build import BUILD as FAKE_DIR, build_fake  # noqa: E402


def _build(tmp_path, libdir):
    exe = tmp_path / "seq_stub"
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", str(exe), str(SRC), f"-L{libdir}", "-lfqd_cuda", f"-Wl,-rpath,{libdir}",
                    "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    return exe


def _check(exe, tmp_path, oracle, env):
    kw = dict(read_len=50, var_len=True, min_len=0, n_frac=0.05, prefix_frac=0.3, sub_frac=0.3, dup_frac=0.5)
    for k, (mode, dist, fmt) in enumerate([("tight", 2, "fastq"), ("loose", 2, "fastq"), ("tail-hamming", 2, "fastq"), ("tail-hamming", 1, "fasta")]):
        seqs = synth.make_reads(9000, seed=50 + k, **kw)
        buf = synth.to_fastq(seqs) if fmt == "fastq" else synth.to_fasta(seqs)
        (tmp_path / "in").write_bytes(buf)
        r = subprocess.run([str(exe), tmp_path / "in", tmp_path / "out", mode, str(dist), fmt, "clusters"], capture_output=True, text=True,
                           env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        ofmt = oracle.FASTQ if fmt == "fastq" else oracle.FASTA
        exp, _, est = oracle.run_oracle(mode, ofmt, buf, dist=dist)
        assert (tmp_path / "out").read_bytes() == exp, (mode, fmt)
        cl, _ = oracle.cluster_text(mode, ofmt, buf, dist=dist)
        assert (tmp_path / "out.clusters").read_bytes() == cl[0], (mode, fmt)
        assert r.stdout == f"{est.total} reads processed, out of which {est.dups} duplicates were removed.\n"


def test_stub_against_the_test_double(tmp_path, oracle):
    build_fake()
    exe = _build(tmp_path, FAKE_DIR)
    _check(exe, tmp_path, oracle, dict(os.environ, LD_LIBRARY_PATH=str(FAKE_DIR)))


@pytest.mark.gpu
def test_stub_against_the_engine(tmp_path, oracle):
    libdir = ROOT / "fastq-dupaway_b200" / "csrc"
    assert (libdir / "libfqd_cuda.so").exists()
    exe = _build(tmp_path, libdir)
    env = dict(os.environ)
    env.pop("LD_LIBRARY_PATH", None)
    _check(exe, tmp_path, oracle, env)

"""Differential test of the command line: random combinations of the reference's options - valid ones, and ones that
break one or several of main()'s rules at once (src/main.cpp:93-172) - go to the reference binary (oracle/_ref, the
unmodified sources) and to the drop-in binary; exit status, stdout and stderr must be equal.  Which rule fires FIRST
when several are broken is the reference's own code, so this pins the order of the checks, not only their wording.
Inputs do not exist, so a combination that passes the argument rules ends in the reference's "Cannot open file" error
in both programs - no GPU is needed for any of it."""
import os
import random
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "fastq-dupaway_b200" / "host" / "fastq-dupaway"


def combo(rng):
    a = []
    if rng.random() < 0.9:
        a += ["-i", rng.choice(["a.fq", "b.fq"])]
    if rng.random() < 0.5:
        a += [rng.choice(["-u", "--input-2"]), rng.choice(["a.fq", "b.fq", "c.fq"])]
    if rng.random() < 0.9:
        a += [rng.choice(["-o", "--output-1"]), "o1.fq"]
    if rng.random() < 0.5:
        a += ["-p", "o2.fq"]
    if rng.random() < 0.4:
        a += ["--format", rng.choice(["fastq", "fasta", "bam", "FASTQ"])]
    if rng.random() < 0.4:
        a += ["--compare-seq", rng.choice(["tight", "loose", "tail-hamming", "fuzzy"])]
    if rng.random() < 0.3:
        a += ["--distance", rng.choice(["0", "1", "5"])]
    if rng.random() < 0.5:
        a += ["--fast"]
    if rng.random() < 0.3:
        a += ["--unordered"]
    if rng.random() < 0.3:
        a += [rng.choice(["-m", "--mem-limit"]), rng.choice(["100", "499", "500", "2048", "10240", "10241"])]
    if rng.random() < 0.3:
        a += ["-v"]
    if rng.random() < 0.3:
        a += ["--write-clusters"]
    return a


def test_random_option_combinations_behave_like_the_reference(tmp_path, oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/fastq-dupaway not built")
    if not EXE.exists():
        subprocess.run(["make", "-s", "-C", str(EXE.parent)], check=True)
    rng = random.Random(2024)
    kinds = set()
    for it in range(400):
        args = combo(rng)
        got = []
        for exe in (oracle.REF_BIN, EXE):
            for f in os.listdir(tmp_path):
                os.remove(tmp_path / f)
            p = subprocess.run([str(exe), *args], capture_output=True, text=True, cwd=tmp_path)
            got.append((p.returncode, p.stdout, p.stderr))
        assert got[0] == got[1], args
        kinds.add(got[0][2].split("\n")[1] if got[0][2].count("\n") > 1 else got[0][2])
    assert len(kinds) >= 8          # the sample really walks through the different rules

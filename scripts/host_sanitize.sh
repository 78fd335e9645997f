#!/bin/bash
# Host I/O layer under ThreadSanitizer and AddressSanitizer + UBSan (CPU only).  Builds two instrumented copies of
# fastq-dupaway_b200/host/io_selftest into /tmp and runs: every archive kind through the block-/member-parallel readers,
# the ring -> runs -> asynchronous writer pipeline, and a few hundred valid and damaged archives through the decoders.
# Round 1 result: 0 reports from either after one fix - ThreadSanitizer found a join object on the writer's stack that a
# worker could still be notifying when the waiter had already returned (io.hpp, OutputFile::write_runs; now shared).
set -e
cd "$(dirname "$0")/../fastq-dupaway_b200/host"
g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -o /tmp/io_selftest_tsan io_selftest.cpp -lz
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -pthread -o /tmp/io_selftest_asan io_selftest.cpp -lz
cd ../..
# the drivers themselves (dup_remover.cpp) over the test double of the engine
python3 -c "import sys; sys.path.insert(0, 'tests/fake_engine'); from build import build_fake; build_fake()"
g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -o /tmp/fqd_cli_tsan fastq-dupaway_b200/host/main.cpp fastq-dupaway_b200/host/options.cpp \
    fastq-dupaway_b200/host/dup_remover.cpp -Ltests/fake_engine/_build -lfqd_cuda -lz
python3 - <<'PY'
import os, random, subprocess, sys, tempfile
sys.path.insert(0, "tests")
from test_host_io import bgzf, deflate_gz, fastq_bytes, members, payloads
tmp = tempfile.mkdtemp(prefix="fqd_san_")
data = fastq_bytes(60000, seed=1)
files = {"single": deflate_gz(data, 6), "multi": members(data, [700000]), "bgzf": bgzf(data)}
small = {"FQD_PINFLATE_CHUNK": "65536", "FQD_GZ_MAX_TASK": "262144", "FQD_GZ_SPAN": "131072"}
reports = 0
for name, blob in files.items():
    p = os.path.join(tmp, name + ".gz")
    open(p, "wb").write(blob)
    r = subprocess.run(["/tmp/io_selftest_tsan", "cat", p, "65536"], capture_output=True, env=dict(os.environ, FQD_IO_THREADS="6", **small))
    assert r.stdout == data, name
    reports += r.stderr.count(b"WARNING: ThreadSanitizer")
plain = os.path.join(tmp, "p.fq")
open(plain, "wb").write(data)
for src, dst in ((plain, "o.fq"), (os.path.join(tmp, "bgzf.gz"), "o.fq.gz")):
    r = subprocess.run(["/tmp/io_selftest_tsan", "filter", src, os.path.join(tmp, dst), "65536"], capture_output=True, env=dict(os.environ, FQD_IO_THREADS="6"))
    assert r.returncode == 0
    reports += r.stderr.count(b"WARNING: ThreadSanitizer")
multi = os.path.join(tmp, "multi.gz"); single = os.path.join(tmp, "single.gz"); bg = os.path.join(tmp, "bgzf.gz")
jobs = [["-i", plain, "-o", tmp + "/c.fq", "--fast"], ["-i", plain, "-o", tmp + "/c.fq", "--compare-seq", "tight", "--write-clusters"],
        ["-i", single, "-u", multi, "-o", tmp + "/c1.fq", "-p", tmp + "/c2.fq.gz", "--fast", "--unordered"],
        ["-i", single, "-u", bg, "-o", tmp + "/c1.fq", "-p", tmp + "/c2.fq", "--compare-seq", "loose"]]
for job in jobs:
    for block in ("262144", "33554432"):
        for policy in ("resident", "discard"):      # whole-input modes: device gather / host gather from the mapped file or the spool
            env = dict(os.environ, LD_LIBRARY_PATH=os.path.abspath("tests/fake_engine/_build"), FQD_IO_THREADS="6", FQD_BLOCK_BYTES=block,
                       FQD_ORDERLY_EXIT="1", FQD_WHOLE_INPUT=policy, **small)
            r = subprocess.run(["/tmp/fqd_cli_tsan", *job], capture_output=True, env=env)
            assert r.returncode == 0, r.stderr[-500:]
            reports += r.stderr.count(b"WARNING: ThreadSanitizer")
print("ThreadSanitizer reports:", reports)
rng = random.Random(99)
P = payloads()
blobs = [deflate_gz(P["mixed"][:2_000_000], 6), members(P["text"][:2_000_000], [300000]), bgzf(P["text"][:2_000_000]),
         deflate_gz(P["text"][:2_000_000], 1, mem=1), deflate_gz(P["runs"][:3_000_000], 6), deflate_gz(P["far"], 9)]
san = 0
for it in range(200):
    blob = bytearray(rng.choice(blobs))
    if it % 5:
        for _ in range(rng.choice([1, 1, 2, 5, 20])):
            pos = rng.randrange(len(blob))
            blob[pos] = blob[pos] ^ (1 << rng.randrange(8)) if rng.random() < 0.5 else rng.getrandbits(8)
        if rng.random() < 0.2:
            blob = blob[:rng.randrange(len(blob))]
    p = os.path.join(tmp, "f.gz")
    open(p, "wb").write(bytes(blob))
    env = dict(os.environ, FQD_IO_THREADS=str(rng.choice([2, 6])), FQD_PINFLATE_CHUNK=str(rng.randrange(4096, 200000)),
               FQD_GZ_MAX_TASK=str(rng.choice([1000, 50000])), FQD_GZ_SPAN=str(rng.choice([64, 30000])), ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run(["/tmp/io_selftest_asan", "cat", p, "65536"], capture_output=True, env=env, timeout=300)
    assert r.returncode in (0, 1), (it, r.returncode)
    san += (b"Sanitizer" in r.stderr) or (b"runtime error" in r.stderr)
print("AddressSanitizer / UBSan reports:", san)
PY

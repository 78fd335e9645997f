"""Host-side ingest / egress of the drop-in binary (fastq-dupaway_b200/host/io.hpp, pargz.hpp), on the CPU.

The reference reads ".gz" through one Boost.Iostreams filter on the main thread (src/file_utils.cpp:59-66) and
writes it the same way (:83-92); here members are inflated in parallel and the output is deflated in parallel.
These tests check that whatever the archive looks like - one member, thousands, BGZF, header magic inside the
compressed bytes, members larger than a task, padding, truncation, corruption - the bytes that reach the engine are
exactly the decompressed file, against Python's gzip module as the independent decoder.
"""
import gzip
import json
import os
import random
import struct
import subprocess
import zlib
from pathlib import Path

import pytest

HOST = Path(__file__).resolve().parent.parent / "fastq-dupaway_b200" / "host"
EXE = HOST / "io_selftest"
MAGIC = b"\x1f\x8b\x08\x00"


@pytest.fixture(scope="module", autouse=True)
def build_selftest():
    subprocess.run(["make", "-s", "-C", str(HOST), "io_selftest"], check=True)
    assert EXE.exists()


def run(args, threads=4, stdin=None, check=True, env=None):
    env = dict(os.environ, FQD_IO_THREADS=str(threads), **(env or {}))
    p = subprocess.run([str(EXE)] + [str(a) for a in args], input=stdin, capture_output=True, env=env, timeout=300)
    if check:
        assert p.returncode == 0, p.stderr.decode()
    return p


def fastq_bytes(n, seed=0, read_len=100):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        seq = "".join(rng.choice("ACGT") for _ in range(read_len))
        out.append(f"@R.{i} 1\n{seq}\n+\n{'I' * read_len}\n")
    return "".join(out).encode()


def members(data, sizes, level=6):
    """gzip members over consecutive slices of data"""
    out, pos, i = [], 0, 0
    while pos < len(data):
        k = sizes[i % len(sizes)]
        out.append(gzip.compress(data[pos:pos + k], compresslevel=level))
        pos += k
        i += 1
    return b"".join(out)


def bgzf(data, block=0xff00, eof_marker=True):
    """BGZF as bgzip writes it: extra subfield BC = member size - 1"""
    out = []
    for pos in list(range(0, len(data), block)) + ([len(data)] if eof_marker else []):
        piece = data[pos:pos + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(piece) + c.flush()
        bsize = 12 + 6 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
                   + body + struct.pack("<II", zlib.crc32(piece), len(piece)))
    return b"".join(out)


def check_file(path, expect, threads=4, block=1 << 16, env=None):
    got = run(["cat", path, block], threads=threads, env=env).stdout
    assert got == expect
    return json.loads(run(["stat", path], threads=threads, env=env).stdout)


@pytest.mark.parametrize("threads", [1, 2, 8])
def test_plain_file(tmp_path, threads):
    data = fastq_bytes(40000, seed=1)          # ~9 MB: the concurrent pread path needs >= 8 MiB per block
    f = tmp_path / "a.fastq"
    f.write_bytes(data)
    assert run(["cat", f, 1 << 16], threads=threads).stdout == data
    assert run(["cat", f, 12 << 20], threads=threads).stdout == data


def test_empty_inputs(tmp_path):
    (tmp_path / "e.fastq").write_bytes(b"")
    (tmp_path / "e.gz").write_bytes(b"")
    (tmp_path / "m.gz").write_bytes(gzip.compress(b""))
    for n in ("e.fastq", "e.gz", "m.gz"):
        assert run(["cat", tmp_path / n]).stdout == b""
    p = run(["cat", tmp_path / "missing.gz"], check=False)
    assert p.returncode == 1 and b"Cannot open file" in p.stderr


@pytest.mark.parametrize("threads", [1, 2, 8])
def test_single_member(tmp_path, threads):
    data = fastq_bytes(20000, seed=2)
    f = tmp_path / "one.fq.gz"
    f.write_bytes(gzip.compress(data))
    st = check_file(f, data, threads)
    assert st["bytes"] == len(data)


@pytest.mark.parametrize("threads", [2, 8])
@pytest.mark.parametrize("sizes", [[1000], [70000, 1, 333333], [3_000_000]])
def test_multi_member(tmp_path, threads, sizes):
    data = fastq_bytes(40000, seed=3)
    f = tmp_path / "multi.fq.gz"
    f.write_bytes(members(data, sizes))
    st = check_file(f, data, threads)
    assert st["parallel"] and st["tasks"] >= 1 and st["serial_members"] == 0


def test_many_tasks_run_ahead(tmp_path):
    # small span is not configurable from outside; 30 MB of poorly compressible data gives > 10 tasks of 2 MiB
    rng = random.Random(5)
    data = bytes(rng.getrandbits(8) for _ in range(1 << 16)) * 16
    data = b"".join(data[i:] + data[:i] for i in range(0, 30))      # ~30 MB, rotated copies
    blob = members(data, [200_000], level=1)
    f = tmp_path / "many.gz"
    f.write_bytes(blob)
    st = check_file(f, data, 8, block=1 << 20)
    assert st["tasks"] >= 2


@pytest.mark.parametrize("threads", [2, 8])
def test_bgzf(tmp_path, threads):
    data = fastq_bytes(60000, seed=4)
    f = tmp_path / "b.fq.gz"
    f.write_bytes(bgzf(data))
    st = check_file(f, data, threads)
    assert st["parallel"] and st["bgzf"] and st["dropped"] == 0 and st["serial_members"] == 0
    assert abs(st["expansion"] - len(data) / f.stat().st_size) < 0.05 * st["expansion"]
    # BGZF followed by ordinary members: the hop list ends, the scan takes over
    tail = fastq_bytes(3000, seed=5)
    f.write_bytes(bgzf(data, eof_marker=False) + members(tail, [50000]))
    check_file(f, data + tail, threads)


@pytest.mark.parametrize("threads", [2, 8])
def test_magic_inside_compressed_bytes(tmp_path, threads):
    """Stored (level 0) members carry the payload verbatim, so the gzip magic inside the payload shows up in the
    compressed stream: false member starts everywhere.  The output must still be exact."""
    rng = random.Random(6)
    payload = bytearray()
    while len(payload) < 6_000_000:
        payload += bytes(rng.getrandbits(8) for _ in range(rng.randrange(10, 3000)))
        payload += MAGIC + bytes(rng.getrandbits(8) for _ in range(6))
        if rng.random() < 0.2:   # a whole valid little member as payload: a false start that even inflates cleanly
            payload += gzip.compress(b"decoy" * rng.randrange(1, 50))
    data = bytes(payload)
    for sizes in ([2_500_000], [100_000, 900_000], [len(data)]):
        f = tmp_path / "decoy.gz"
        f.write_bytes(members(data, sizes, level=0))
        st = check_file(f, data, threads)
        assert st["parallel"]
    # one huge stored member with decoys followed by small real members
    tail = fastq_bytes(5000, seed=7)
    f.write_bytes(gzip.compress(data, compresslevel=0) + members(tail, [20000]))
    check_file(f, data + tail, threads)


def test_member_larger_than_a_task_goes_serial(tmp_path):
    rng = random.Random(8)
    big = bytes(rng.getrandbits(8) for _ in range(1 << 20)) * 40       # 40 MiB, compresses badly at level 1... stored
    tail = fastq_bytes(5000, seed=9)
    f = tmp_path / "big.gz"
    f.write_bytes(gzip.compress(big, compresslevel=0) + members(tail, [100000]) + gzip.compress(big[:1 << 20], compresslevel=0))
    st = check_file(f, big + tail + big[:1 << 20], 4, block=1 << 20)
    assert st["serial_members"] >= 1 and st["tasks"] >= 1


def test_padding_truncation_corruption(tmp_path):
    data = fastq_bytes(20000, seed=10)
    blob = members(data, [300000])
    f = tmp_path / "x.gz"
    # zero padding after the last member is ignored (as gzip(1) does)
    f.write_bytes(blob + b"\0" * 1000)
    check_file(f, data)
    # garbage after the last member: error, like the serial zlib path
    f.write_bytes(blob + b"garbage-garbage-garbage-garbage")
    for t in (1, 4):
        p = run(["cat", f], threads=t, check=False)
        assert p.returncode == 1 and b"gzip error" in p.stderr
    # cut in the middle of a member: the bytes before the cut are delivered, then end of input (both paths agree)
    f.write_bytes(blob[:len(blob) // 2])
    a = run(["cat", f], threads=1).stdout
    b = run(["cat", f], threads=4).stdout
    assert data.startswith(b) and len(b) > 0
    assert data.startswith(a)
    # a flipped bit in the middle: CRC / inflate error
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0x10
    f.write_bytes(bytes(bad))
    p = run(["cat", f], threads=4, check=False)
    assert p.returncode == 1 and b"gzip error" in p.stderr
    # not gzip at all
    f.write_bytes(b"@r\nACGT\n+\nIIII\n" * 10)
    p = run(["cat", f], threads=4, check=False)
    assert p.returncode == 1 and b"gzip error" in p.stderr


@pytest.mark.parametrize("threads", [1, 2, 8])
def test_gz_output(tmp_path, threads):
    data = fastq_bytes(30000, seed=11)           # ~6.6 MB: several 1 MiB pieces
    out = tmp_path / "o.fq.gz"
    run(["put", out], threads=threads, stdin=data)
    assert gzip.decompress(out.read_bytes()) == data
    assert subprocess.run(["gzip", "-t", str(out)]).returncode == 0
    # and back in through the parallel reader
    check_file(out, data, threads)
    # FQD_GZ_LEVEL: faster, larger
    size6 = out.stat().st_size
    run(["put", out], threads=threads, stdin=data, env={"FQD_GZ_LEVEL": "1"})
    assert gzip.decompress(out.read_bytes()) == data and out.stat().st_size > size6
    # empty output is still a valid archive
    run(["put", out], threads=threads, stdin=b"")
    assert gzip.decompress(out.read_bytes()) == b""
    # plain output
    plain = tmp_path / "o.fq"
    run(["put", plain], threads=threads, stdin=data)
    assert plain.read_bytes() == data


def every_third_dropped(data):
    lines = data.split(b"\n")[:-1]
    recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines), 4)]
    return b"".join(r for i, r in enumerate(recs) if i % 3 != 1)


@pytest.mark.parametrize("threads", [1, 2, 8])
def test_driver_pipeline_with_stand_in_verdict(tmp_path, threads):
    """BlockReader ring -> MateStream (tail carry) -> survivor runs -> AsyncWriter -> OutputFile, the way
    dup_remover.cpp strings them together; blocks are recycled behind the asynchronous writes.  Tiny blocks make
    every record straddle, a large block takes the concurrent pwrite path (>= 8 MiB of survivors per chunk)."""
    data = fastq_bytes(80000, seed=12)           # 17.8 MB
    expect = every_third_dropped(data)
    src = tmp_path / "in.fq"
    src.write_bytes(data)
    gzsrc = tmp_path / "in.fq.gz"
    gzsrc.write_bytes(members(data, [1_000_000]))
    for block in (4096, 1 << 16, 32 << 20):
        out = tmp_path / "out.fq"
        st = json.loads(run(["filter", src, out, block], threads=threads).stdout)
        assert st["records"] == 80000
        assert out.read_bytes() == expect, block
    outgz = tmp_path / "out.fq.gz"
    run(["filter", gzsrc, outgz, 1 << 20], threads=threads)
    assert gzip.decompress(outgz.read_bytes()) == expect


# ---- one member, many threads (pinflate.hpp) ---------------------------------------------------------------------
# small files are cut into many pieces: 16 KiB chunks, every member above 64 KiB goes to the block-parallel decoder
SMALL = {"FQD_PINFLATE_CHUNK": str(16 << 10), "FQD_GZ_MAX_TASK": str(64 << 10), "FQD_GZ_SPAN": str(32 << 10)}


def deflate_gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    c = zlib.compressobj(level, zlib.DEFLATED, 31, mem, strategy)
    return c.compress(data) + c.flush()


def payloads():
    rng = random.Random(21)
    text = fastq_bytes(30000, seed=20)                                   # 6.6 MB of FASTQ
    noise = bytes(rng.getrandbits(8) for _ in range(1 << 20))            # incompressible: stored blocks
    runs = b"".join(bytes([rng.randrange(4)]) * rng.randrange(1, 70000) for _ in range(200))   # long matches, distance 1
    far = (noise[:30000] + b"x" * 2000) * 60                             # matches at distances close to 32 KiB
    # lines of unique bytes + a constant: no long-lived references into the unknown window, so a chunk switches from
    # 16-bit symbols to plain bytes early
    unique = b"".join(bytes(rng.getrandbits(8) | 0x80 for _ in range(40)) + b"A" * 64 + b"\n" for _ in range(30000))
    return {"text": text, "noise": noise, "runs": runs, "far": far, "unique": unique,
            "mixed": text[:2_000_000] + noise + runs + far + text[2_000_000:]}


@pytest.mark.parametrize("threads", [2, 8])
@pytest.mark.parametrize("name", ["text", "noise", "runs", "far", "unique", "mixed"])
def test_single_member_block_parallel(tmp_path, threads, name):
    data = payloads()[name]
    f = tmp_path / "one.gz"
    for level, strategy in ((6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                            (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (0, zlib.Z_DEFAULT_STRATEGY)):
        f.write_bytes(deflate_gz(data, level, strategy))
        st = check_file(f, data, threads, env=SMALL)
        if f.stat().st_size > (64 << 10) + 100:        # smaller archives are one ordinary task
            assert st["serial_members"] == 1 and st["member_chunks"] >= 1, (level, strategy, st)
    # small blocks (memLevel 1): thousands of block boundaries
    f.write_bytes(deflate_gz(data, 6, mem=1))
    st = check_file(f, data, threads, env=SMALL)
    assert st["member_chunks"] > 4 or name in ("runs", "noise", "far")
    if name == "unique":
        f.write_bytes(deflate_gz(data, 6))
        st = check_file(f, data, threads, env=dict(SMALL, FQD_PINFLATE_CHUNK=str(256 << 10)))
        assert st["direct_bytes"] > 500_000                     # chunks went on in plain bytes after their first blocks


def test_single_member_block_parallel_default_sizes(tmp_path):
    """default sizes: 1 MiB chunks, members above 8 MiB of compressed data"""
    data = fastq_bytes(30000, seed=22) * 8            # 53 MB, ~10 MB compressed at level 1
    f = tmp_path / "big.fq.gz"
    f.write_bytes(deflate_gz(data, 1))
    assert f.stat().st_size > (8 << 20)
    st = check_file(f, data, 8, block=4 << 20)
    assert st["serial_members"] == 1 and st["member_chunks"] >= 4
    # the measured expansion (sizes the device tables for .gz input) is the true one once the file is through
    assert abs(st["expansion"] - len(data) / f.stat().st_size) < 0.05 * st["expansion"]
    # the same through the serial zlib path
    assert run(["cat", f, 4 << 20], threads=8, env={"FQD_PINFLATE": "0"}).stdout == data


def test_single_member_header_fields_and_neighbours(tmp_path):
    import io
    data = payloads()["text"]
    buf = io.BytesIO()
    with gzip.GzipFile(filename="reads_with_a_name.fastq", mode="wb", fileobj=buf, compresslevel=6, mtime=12345) as g:
        g.write(data)
    named = buf.getvalue()
    assert named[3] & 8                                # FNAME present
    extra = bytearray(deflate_gz(data))
    extra[3] |= 4                                      # FEXTRA: splice a subfield in after the fixed header
    extra = bytes(extra[:10]) + struct.pack("<H", 8) + b"ZZ" + struct.pack("<H", 4) + b"abcd" + bytes(extra[10:])
    small = fastq_bytes(300, seed=23)
    f = tmp_path / "n.gz"
    for blob, expect in ((named, data), (extra, data),
                         (members(small, [10000]) + named + members(small, [7000]) + extra, small + data + small + data)):
        f.write_bytes(blob)
        check_file(f, expect, 4, env=SMALL)


def test_single_member_truncated_and_corrupt(tmp_path):
    data = payloads()["text"]
    blob = deflate_gz(data)
    f = tmp_path / "t.gz"
    f.write_bytes(blob[:len(blob) * 2 // 3])
    a = run(["cat", f], threads=1).stdout                      # serial zlib
    b = run(["cat", f], threads=4, env=SMALL).stdout           # block-parallel
    assert a == b and data.startswith(b) and len(b) > len(data) // 2
    # trailer cut off: all the data, no error (as the serial path)
    f.write_bytes(blob[:-5])
    assert run(["cat", f], threads=4, env=SMALL).stdout == data
    for pos in (len(blob) // 3, len(blob) - 6, len(blob) - 2):    # body, CRC-32, ISIZE
        bad = bytearray(blob)
        bad[pos] ^= 0x04
        f.write_bytes(bytes(bad))
        p = run(["cat", f], threads=4, env=SMALL, check=False)
        assert p.returncode == 1 and b"gzip error" in p.stderr, pos


def test_damaged_archives_never_crash_and_agree_with_the_serial_path(tmp_path):
    """Bit flips, overwritten bytes and cuts in single-member, multi-member, BGZF and small-block archives: the
    parallel readers either report `gzip error` or deliver exactly what the serial zlib path (FQD_IO_THREADS=1)
    delivers for the same file - never a crash, a hang, or bytes of their own."""
    rng = random.Random(77)
    text = payloads()["mixed"][:1_500_000]
    blobs = [deflate_gz(text, 6), members(text, [300000]), bgzf(text), deflate_gz(text, 1, mem=1)]
    f = tmp_path / "f.gz"
    outcomes = {"error": 0, "ok": 0}
    for it in range(36):
        blob = bytearray(blobs[it % 4])
        for _ in range(rng.choice([1, 1, 2, 5])):
            pos = rng.randrange(len(blob))
            if rng.random() < 0.5:
                blob[pos] ^= 1 << rng.randrange(8)
            else:
                blob[pos] = rng.getrandbits(8)
        if rng.random() < 0.3:
            blob = blob[:rng.randrange(len(blob))]
        f.write_bytes(bytes(blob))
        env = {"FQD_PINFLATE_CHUNK": str(rng.randrange(4096, 200000)), "FQD_GZ_MAX_TASK": str(rng.choice([1000, 50000])),
               "FQD_GZ_SPAN": str(rng.choice([64, 30000]))}
        p = run(["cat", f, 65536], threads=rng.choice([2, 8]), env=env, check=False)
        assert p.returncode in (0, 1), (it, p.returncode, p.stderr[-200:])
        if p.returncode == 1:
            assert b"gzip error" in p.stderr
            outcomes["error"] += 1
        else:
            serial = run(["cat", f, 65536], threads=1, check=False)
            assert serial.returncode == 0 and serial.stdout == p.stdout, it
            outcomes["ok"] += 1
    assert outcomes["error"] > 10


def test_enormous_expansion_stays_within_the_run_ahead_budget(tmp_path):
    """512 MiB of zeros are 0.5 MB of deflate: every 64 KiB chunk inflates to 64 MiB.  The reader must not keep
    2 x threads of those in flight; FQD_IO_AHEAD_MB bounds it (peak RSS checked), and the bytes are still right."""
    import resource
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    z = bytes(1 << 24)
    blob = b"".join(c.compress(z) for _ in range(32)) + c.flush()
    f = tmp_path / "zeros.gz"
    f.write_bytes(blob)
    env = {"FQD_PINFLATE_CHUNK": str(64 << 10), "FQD_GZ_MAX_TASK": str(64 << 10), "FQD_IO_AHEAD_MB": "64"}
    before = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    st = json.loads(run(["stat", f], threads=8, env=env).stdout)
    peak_mb = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss // 1024
    assert st["bytes"] == 32 << 24 and st["member_chunks"] >= 4
    if peak_mb * 1024 > before:          # ru_maxrss is the maximum over all children so far
        assert peak_mb < 1200, peak_mb   # 8 chunks in flight would hold more than 2 GB of symbols
    p = run(["cat", f, 1 << 20], threads=8, env=env)
    assert len(p.stdout) == 32 << 24 and p.stdout.count(0) == len(p.stdout)

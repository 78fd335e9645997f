"""GPU test of the multi-GPU sequence-based mode (key-range sharding, fastq-dupaway_b200/sharded_seq.py): three ranks,
all on cuda:0, exchanging through gloo on the host (NCCL refuses several ranks on one device).  The concatenation of
the ranks' outputs must be byte-identical to the oracle's output for the whole input, in every --compare-seq mode;
the inputs are low-complexity short reads, so that prefix and Hamming clusters straddle the range boundaries."""
import importlib
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, mode, dval, fmt_name, slices, result_dir, use_peer=False):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fqd = importlib.import_module("fastq-dupaway_b200")
    sh = importlib.import_module("fastq-dupaway_b200.sharded_seq")
    torch.cuda.set_device(0)
    fmt = fqd.FORMAT_FASTQ if fmt_name == "fastq" else fqd.FORMAT_FASTA
    mine = slices[rank]
    paired = len(mine) == 2
    ops = sh.GpuRangeOps(fqd, mode, fmt, paired, dval, 64, 40000, 40000, 0, seg_bytes=1 << 18)
    for m, b in enumerate(mine):
        if b:
            d = fqd.DeviceBuffer(len(b), 0)
            d.upload(b)
            ops.append(m, d.ptr, len(b))
            torch.cuda.synchronize()
            d.free()
    n_local_expected = mine[0].count(b"\n") // (4 if fmt_name == "fastq" else 2)
    st0 = None
    peers = None
    if use_peer:          # records over mapped peer memory, parsed in place where they land (fqd_adopt_device)
        px = importlib.import_module("fastq-dupaway_b200.peer")
        peers = [px.PeerExchange(fqd, dist, rank, world, 0, 1 << 20) for _ in mine]
    owned, kept, dups = sh.dedup_ranges(ops, dist, rank, world, n_samples=256, via_cpu=True, peer=peers)
    st0 = ops.origin.stats()
    assert ops.origin.partition_sample(1)[1] == n_local_expected, (
        "origin parse", rank, ops.origin.partition_sample(1)[1], n_local_expected, st0.err, st0.err_record, st0.err_char, st0.err_mate,
        [len(b) for b in mine])
    for m in range(len(mine)):
        (Path(result_dir) / f"out_{rank}_{m}.bin").write_bytes(ops.output(m) if owned else b"")
    (Path(result_dir) / f"cnt_{rank}.txt").write_text(f"{owned} {kept} {dups}")
    dist.barrier()
    ops.close()
    if peers:
        for p_ in peers:
            p_.close()
    dist.destroy_process_group()


def _run(tmp_path, oracle, mode, dval, fmt_name, b1, b2, world, cuts, use_peer=False):
    """cuts: record-aligned byte offsets that split the input into `world` contiguous slices"""
    bufs = [b1] if b2 is None else [b1, b2]
    slices = []
    for r in range(world):
        slices.append([b[c[r]: c[r + 1]] for b, c in zip(bufs, cuts)])
    port = 29900 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, mode, dval, fmt_name, slices, str(tmp_path), use_peer), nprocs=world, join=True)
    fmt = oracle.FASTQ if fmt_name == "fastq" else oracle.FASTA
    e1, e2, est = oracle.run_oracle(mode, fmt, b1, b2, dist=dval)
    got = [b"".join((tmp_path / f"out_{r}_{m}.bin").read_bytes() for r in range(world)) for m in range(len(bufs))]
    cnt = [tuple(int(x) for x in (tmp_path / f"cnt_{r}.txt").read_text().split()) for r in range(world)]
    assert sum(c[0] for c in cnt) == est.total
    assert sum(c[2] for c in cnt) == est.dups
    assert got[0] == e1
    if b2 is not None:
        assert got[1] == e2
    return cnt


def _cuts(recs, world, fracs=None):
    """byte offsets of `world` contiguous slices of the record list"""
    n = len(recs)
    bounds = [0] + [int(n * f) for f in (fracs or [(k + 1) / world for k in range(world - 1)])] + [n]
    offs = [0]
    for r in recs:
        offs.append(offs[-1] + len(r))
    return [offs[b] for b in bounds]


MODES = [("tight", 2), ("loose", 2), ("tail-hamming", 1), ("tail-hamming", 2)]


@pytest.mark.parametrize("mode,dval", MODES)
def test_three_ranks_single_end(tmp_path, oracle, mode, dval):
    # 2-letter alphabet, 6..24 bases: thousands of prefix / Hamming relations, many across the two range boundaries
    seqs = synth.make_reads(9000, seed=61, read_len=24, var_len=True, min_len=6, dup_frac=0.5, prefix_frac=0.35, sub_frac=0.35,
                            alphabet=b"AC")
    recs = [synth.to_fastq([s], ids=[b"@r.%d" % i]) for i, s in enumerate(seqs)]
    b1 = b"".join(recs)
    cnt = _run(tmp_path, oracle, mode, dval, "fastq", b1, None, 3, [_cuts(recs, 3, [0.2, 0.7])])
    assert all(c[0] > 0 for c in cnt)


@pytest.mark.parametrize("mode,dval", MODES)
def test_three_ranks_paired(tmp_path, oracle, mode, dval):
    s1, s2 = synth.make_pair(6000, seed=62, read_len=16, var_len=True, min_len=4, dup_frac=0.5, prefix_frac=0.3, sub_frac=0.3, alphabet=b"AC")
    r1 = [synth.to_fasta([s], ids=[b">p.%d 1" % i]) for i, s in enumerate(s1)]
    r2 = [synth.to_fasta([s], ids=[b">p.%d 2" % i]) for i, s in enumerate(s2)]
    _run(tmp_path, oracle, mode, dval, "fasta", b"".join(r1), b"".join(r2), 3, [_cuts(r1, 3), _cuts(r2, 3)])


def test_identical_reads_and_an_empty_range(tmp_path, oracle):
    # every read identical: one key range gets everything, the others are empty and only pass the boundary state on
    recs = [synth.to_fastq([b"ACGTACGTACGTACGTACGTAAAA"], ids=[b"@x.%d" % i]) for i in range(500)]
    cnt = _run(tmp_path, oracle, "tail-hamming", 2, "fastq", b"".join(recs), None, 3, [_cuts(recs, 3)])
    assert sorted(c[0] for c in cnt) == [0, 0, 500]


@pytest.mark.parametrize("mode,dval", [("loose", 2), ("tail-hamming", 2)])
def test_three_ranks_paired_over_peer_memory(tmp_path, oracle, mode, dval):
    """Same as above with the production exchange: CUDA IPC receive buffers, copy-engine writes, records parsed in place."""
    s1, s2 = synth.make_pair(6000, seed=63, read_len=20, var_len=True, min_len=4, dup_frac=0.5, prefix_frac=0.3, sub_frac=0.3, alphabet=b"AC")
    r1 = [synth.to_fastq([s], ids=[b"@p.%d 1" % i]) for i, s in enumerate(s1)]
    r2 = [synth.to_fastq([s], ids=[b"@p.%d 2" % i]) for i, s in enumerate(s2)]
    _run(tmp_path, oracle, mode, dval, "fastq", b"".join(r1), b"".join(r2), 3, [_cuts(r1, 3), _cuts(r2, 3)], use_peer=True)

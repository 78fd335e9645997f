"""GPU test of the multi-GPU --fast --unordered path (tag-range sharding, fastq-dupaway_b200/sharded_unordered.py):
three ranks, all on cuda:0, exchanging through gloo on the host (NCCL refuses several ranks on one device).  The
concatenation of the ranks' outputs must be byte-identical to the oracle's output for the whole input; the cases are the
ones tests/test_sharded_unordered_cpu.py runs against the pure-Python stand-in - hundreds of tiny tag multisets that put
the reference's end-of-stream rule on, before and after every range boundary, and shuffled FASTQ files
with deletions, duplicate pairs, a repeated tag and a byte outside {A,C,G,T,N}."""
import importlib
import os
import pickle
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist

import test_sharded_unordered_cpu as proto

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, case_file, result_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fqd = importlib.import_module("fastq-dupaway_b200")
    sh = importlib.import_module("fastq-dupaway_b200.sharded_unordered")
    torch.cuda.set_device(0)
    cases = pickle.loads(Path(case_file).read_bytes())
    results = []
    ops = None
    for lpr, slices, n_samples in cases:
        if ops is None:
            ops = sh.GpuTagRangeOps(fqd, fqd.FORMAT_FASTQ if lpr == 4 else fqd.FORMAT_FASTA, 64, 20000, 20000, 0, seg_bytes=1 << 18)
        else:
            ops.reset()
        for m, b in enumerate(slices[rank]):
            if b:
                d = fqd.DeviceBuffer(len(b), 0)
                d.upload(b)
                ops.append(m, d.ptr, len(b))
                torch.cuda.synchronize()
                d.free()
        res = sh.dedup_tag_ranges(ops, dist, rank, world, n_samples=n_samples, via_cpu=True)
        outs = (b"", b"") if res.err == sh.ERR_EMPTY else (ops.output(0), ops.output(1))
        if res.err == sh.ERR_BAD_BASE:
            st = ops.range.stats()
            assert st.err in (0, 6)
        results.append((res, outs))
    (Path(result_dir) / f"res_{rank}.pkl").write_bytes(pickle.dumps(results))
    dist.barrier()
    if ops is not None:
        ops.close()
    dist.destroy_process_group()


def test_stop_rule_over_every_boundary(tmp_path, oracle):
    cases, wholes = proto.tiny_cases(150, seed=11, world=3)
    got = proto.run_cases(tmp_path, cases, 3, worker=_worker, port_base=33900)
    for whole, g in zip(wholes, got):
        proto.check_against_oracle(oracle, oracle.FASTA, whole, g)


def test_shuffled_files_with_deletions_duplicates_and_a_bad_base(tmp_path, oracle):
    cases, wholes = proto.big_cases(12, n=3000)
    got = proto.run_cases(tmp_path, cases, 3, worker=_worker, port_base=33900)
    for whole, g in zip(wholes, got):
        proto.check_against_oracle(oracle, oracle.FASTQ, whole, g)
    assert got[2][0].err == 6


def test_one_file_empty(tmp_path, oracle):
    cases = [(2, [(b">X.1\nACGT\n", b""), (b"", b""), (b">X.2\nACGT\n", b"")], 4)]
    got = proto.run_cases(tmp_path, cases, 3, worker=_worker, port_base=33900)
    assert got[0][0].err == 3


def test_two_ranks_tags_that_differ_behind_the_splitter_words(tmp_path, oracle):
    """every tag shares the 16 bytes the splitters look at: one range owns everything, the other one is empty"""
    import numpy as np
    rng = np.random.default_rng(5)
    recs = [[b">LONGPREFIX_LONGPREFIX_%04d %d\n%s\n" % (int(rng.integers(0, 300)), mate + 1, [b"ACGT", b"AAAA", b"ACGA"][int(rng.integers(0, 3))])
             for _ in range(500)] for mate in range(2)]
    sl = [proto.slice_records(r, 2, rng) for r in recs]
    cases = [(2, [(sl[0][r], sl[1][r]) for r in range(2)], 32)]
    got = proto.run_cases(tmp_path, cases, 2, worker=_worker, port_base=33900)
    proto.check_against_oracle(oracle, oracle.FASTA, (b"".join(recs[0]), b"".join(recs[1])), got[0])


def test_one_rank_is_the_whole_job(tmp_path, oracle):
    cases, wholes = proto.tiny_cases(60, seed=13, world=1)
    got = proto.run_cases(tmp_path, cases, 1, worker=_worker, port_base=33900)
    for whole, g in zip(wholes, got):
        proto.check_against_oracle(oracle, oracle.FASTA, whole, g)


def test_stage_calls_out_of_order_are_refused(fqd):
    eng = fqd.Engine("fast", fqd.FORMAT_FASTA, True, True, 2, 64, 1000, 1 << 18, 0, 0)
    with pytest.raises(fqd.FqdError):
        eng.unordered_join(0, 0, 0, 0)              # before fqd_unordered_prepare
    with pytest.raises(fqd.FqdError):
        eng.unordered_enter(0, 0)
    eng.append(0, b">X.1\nACGT\n>X.2\nACGT\n")
    eng.append(1, b">X.2\nACGT\n")
    assert eng.unordered_prepare() == (2, 1)
    with pytest.raises(fqd.FqdError):
        eng.unordered_prepare()                     # twice
    with pytest.raises(fqd.FqdError):
        eng.unordered_enter(0, 3)                   # outside the list
    assert eng.unordered_enter(0, 2) == 1 and eng.unordered_enter(1, 1) == 2
    with pytest.raises(fqd.FqdError):
        eng.unordered_apply(0, 0, False)            # before fqd_unordered_join
    eng.close()
    plain = fqd.Engine("tight", fqd.FORMAT_FASTA, False, False, 2, 64, 1000, 1 << 18, 0, 0)
    with pytest.raises(fqd.FqdError):
        plain.unordered_prepare()
    plain.close()

"""The CPU differential suites of test_cli_host_logic.py / test_cli_whole_input_host_logic.py once more, but with the REAL
libfqd_cuda.so behind the binary - i.e. the engine itself against the reference binaries (oracle/_ref travels to the GPU
box) on malformed records swept over the file, input cut inside the last record, double faults, odd and unusual files,
in every mode.  Every run of the real binary pays 1 - 3 s of CUDA start-up, so under `-m gpu` the sweeps visit two
positions each (one early, one in the last block) and fewer flag sets - ~80 runs; the CPU suites over the test double
visit every position, and FQD_DIFF_FULL=1 restores all ~800 runs here as well (over 30 min on a B200 box).
"""
import importlib
import os
import subprocess

import pytest

import test_cli_host_logic as fast_suite
import test_cli_whole_input_host_logic as whole_suite

pytestmark = pytest.mark.gpu


def _real_run(*args, env=None):
    e = dict(os.environ, FQD_IO_THREADS="4")
    e.pop("LD_LIBRARY_PATH", None)
    e.update(env or {})
    return subprocess.run([str(fast_suite.EXE), *map(str, args)], capture_output=True, text=True, env=e, timeout=600)


@pytest.fixture(autouse=True)
def real_engine(monkeypatch):
    """Point both suites at the product library: their run() helpers stop preloading the test double."""
    monkeypatch.setattr(fast_suite, "run", _real_run)
    monkeypatch.setattr(whole_suite, "run", _real_run)
    monkeypatch.setattr(whole_suite, "FAKE_DIR", "/nonexistent")      # the two tests that build their own environment
    if not os.environ.get("FQD_DIFF_FULL"):
        monkeypatch.setattr(fast_suite, "SWEEP_STRIDE", 7)
        monkeypatch.setattr(whole_suite, "THIN", True)


@pytest.mark.parametrize("kind", ["start", "length", "base"])
def test_fast_malformed_record_everywhere(tmp_path, oracle, kind):
    fast_suite.test_malformed_record_at_every_position_matches_the_reference_binary(tmp_path, oracle, kind)


@pytest.mark.parametrize("kind", ["start", "length", "base"])
@pytest.mark.parametrize("bad_mate", [0, 1])
def test_fast_malformed_record_paired(tmp_path, oracle, kind, bad_mate):
    fast_suite.test_malformed_record_in_paired_input_matches_the_reference_binary(tmp_path, oracle, kind, bad_mate)


def test_fast_input_cut_anywhere(tmp_path, oracle):
    fast_suite.test_input_cut_anywhere_in_the_last_record_matches_the_reference_binary(tmp_path, oracle)


def test_fast_double_faults(tmp_path, oracle):
    fast_suite.test_two_malformed_records_report_the_one_the_reference_meets_first(tmp_path, oracle)


@pytest.mark.parametrize("kind", ["start", "base", "empty", "lower"])
def test_fast_malformed_fasta(tmp_path, oracle, kind):
    fast_suite.test_malformed_fasta_record_at_every_position_matches_the_reference_binary(tmp_path, oracle, kind)


@pytest.mark.parametrize("mode,dist", whole_suite.MODES)
@pytest.mark.parametrize("paired", [False, True])
def test_sequence_modes(tmp_path, oracle, mode, dist, paired):
    whole_suite.test_sequence_modes_against_oracle_and_stable_reference(tmp_path, oracle, mode, dist, paired)


@pytest.mark.parametrize("mode,unordered", [("tight", False), ("tail-hamming", False), ("fast", True)])
def test_whole_input_malformed(tmp_path, oracle, mode, unordered):
    whole_suite.test_malformed_record_in_whole_input_modes_matches_the_reference_binary(tmp_path, oracle, mode, unordered)


def test_odd_inputs(tmp_path, oracle):
    whole_suite.test_odd_inputs_in_every_mode_match_the_reference_binaries(tmp_path, oracle)


def test_unusual_files(tmp_path, oracle):
    whole_suite.test_unusual_files_in_every_mode_match_the_reference_binaries(tmp_path, oracle)


# ---- the same jobs with nothing but key rows resident (FQD_WHOLE_INPUT=discard: inputs larger than device memory) ----
def _discard_run(*args, env=None):
    res = _real_run(*args, env=dict(env or {}, FQD_WHOLE_INPUT="discard", FQD_TRACE="1"))
    if res.returncode == 0:
        assert "raw input not kept on the device" in res.stderr and "gathered on the host" in res.stderr
    res.stderr = "".join(l + "\n" for l in res.stderr.splitlines() if not l.startswith(("[host-trace]", "[fqd trace]")))
    return res


@pytest.mark.parametrize("mode,dist", whole_suite.MODES)
@pytest.mark.parametrize("paired", [False, True])
def test_sequence_modes_discarded_input(tmp_path, oracle, monkeypatch, mode, dist, paired):
    """.gz input spooled, plain input mapped, output records and cluster files gathered on the host from the engine's
    (offset, length) lists: bytes equal to the oracle and to the stable-sort reference binary."""
    monkeypatch.setattr(whole_suite, "run", _discard_run)
    whole_suite.test_sequence_modes_against_oracle_and_stable_reference(tmp_path, oracle, mode, dist, paired)


def test_unordered_discarded_input(tmp_path, oracle, monkeypatch):
    monkeypatch.setattr(whole_suite, "run", _discard_run)
    whole_suite.test_unordered_with_long_tags_restarts_and_matches_the_reference(tmp_path, oracle)


@pytest.mark.parametrize("mode,unordered", [("tight", False), ("fast", True)])
def test_whole_input_malformed_discarded_input(tmp_path, oracle, monkeypatch, mode, unordered):
    monkeypatch.setattr(whole_suite, "run", _discard_run)
    whole_suite.test_malformed_record_in_whole_input_modes_matches_the_reference_binary(tmp_path, oracle, mode, unordered)


# ---- pipes: no size, read once - spooled by default, the spool is what a restart reads ----------------------------------
def test_fifo_in_and_out_of_a_sequence_mode(tmp_path, oracle):
    whole_suite.test_fifo_input_and_fifo_output_in_a_sequence_mode(tmp_path, oracle)


def test_pipe_input_is_spooled_and_survives_a_restart(tmp_path, oracle, monkeypatch):
    import test_cli_whole_input_discard_host_logic as discard_suite
    monkeypatch.setattr(discard_suite, "run", _real_run)
    monkeypatch.setenv("FQD_WHOLE_INPUT", "unset-by-the-test")
    discard_suite.test_pipe_input_survives_a_restart(tmp_path, oracle, monkeypatch, "tail-hamming")

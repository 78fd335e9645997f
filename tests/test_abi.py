"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol of include/fqd.h."""
import ctypes


def test_library_exports_every_declared_symbol(fqd):
    lib = fqd.load_library()
    names = fqd.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fqd.h but not exported"


def test_abi_version(fqd):
    lib = fqd.load_library()
    assert lib.fqd_abi_version() == fqd.FQD_ABI_VERSION


def test_struct_layouts_match_header(fqd):
    # sizes as the C compiler lays them out (checked against a tiny C program would need a compiler run;
    # the header uses only fixed-width fields in natural alignment, so ctypes' layout is the C layout)
    assert ctypes.sizeof(fqd.Config) == 64
    assert ctypes.sizeof(fqd.Stats) == 48
    assert ctypes.sizeof(fqd.ChunkResult) == 64


def test_create_fails_loudly_without_gpu(fqd):
    import torch
    if torch.cuda.is_available():
        return
    try:
        fqd.Engine("fast")
    except fqd.FqdError as e:
        assert e.code == 2      # FQD_ERR_CUDA - never a CPU fallback
    else:
        raise AssertionError("fqd_create succeeded without a CUDA device")


def test_binary_fails_loudly_without_a_gpu(tmp_path):
    """The drop-in binary has no CPU path either: without a CUDA device it stops with the reference's error banner
    and exit status 1 instead of producing output some other way."""
    import subprocess
    from pathlib import Path
    import pytest
    import torch
    ROOT = Path(__file__).resolve().parent.parent
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = ROOT / "fastq-dupaway_b200" / "host" / "fastq-dupaway"
    if not exe.exists():
        pytest.skip("host binary not built")
    inp = tmp_path / "in.fq"
    inp.write_bytes(b"@r1\nACGT\n+\nIIII\n")
    for extra in (["--fast"], []):
        res = subprocess.run([str(exe), "-i", str(inp), "-o", str(tmp_path / "out.fq"), *extra], capture_output=True, text=True)
        assert res.returncode == 1
        assert res.stderr.startswith("An error occured during fastq-dupaway execution:")
        assert "CUDA" in res.stderr


def test_binding_loads_the_product_library_only(fqd):
    """The ctypes binding opens fastq-dupaway_b200/csrc/libfqd_cuda.so by path - never the test double of
    tests/fake_engine, whatever LD_LIBRARY_PATH says."""
    from pathlib import Path
    lib = fqd.load_library()
    assert Path(lib._name).resolve() == (Path(fqd.__file__).resolve().parent / "csrc" / "libfqd_cuda.so")

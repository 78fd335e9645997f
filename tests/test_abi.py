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

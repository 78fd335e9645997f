import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fqd():
    """The product binding (ctypes over libfqd_cuda.so)."""
    return importlib.import_module("fastq-dupaway_b200")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/) - test infrastructure only."""
    sys.path.insert(0, str(ROOT / "oracle"))
    mod = importlib.import_module("oracle")
    mod.lib()
    return mod


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"

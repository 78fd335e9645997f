"""Builds the test double of the C ABI (fake_fqd.cpp + the oracle's C functions) into tests/fake_engine/_build."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
BUILD = HERE / "_build"
LIB = BUILD / "libfqd_cuda.so"


def build_fake() -> Path:
    BUILD.mkdir(exist_ok=True)
    obj = BUILD / "fqd_oracle.o"
    subprocess.run(["gcc", "-O2", "-fPIC", "-c", str(ROOT / "oracle" / "fqd_oracle.c"), "-o", str(obj)], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", str(LIB), str(HERE / "fake_fqd.cpp"), str(obj)], check=True)
    return LIB

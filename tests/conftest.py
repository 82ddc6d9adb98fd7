import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "heif-decoder-lib_b200")
for p in (PKG_DIR, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

STREAMS = os.path.join(ROOT, "tests", "golden", "streams")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """The product libraries are built in-tree by __graft_entry__.build(); build them when a test
    run starts from a clean checkout."""
    need = [os.path.join(PKG_DIR, "libheifcuda_host.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", os.path.join(PKG_DIR, "csrc"), "-j8", "../libheifcuda_host.so"],
                              stdout=subprocess.DEVNULL)


_ensure_built()


@pytest.fixture(scope="session")
def streams_dir():
    return STREAMS


def read_stream(name):
    with open(os.path.join(STREAMS, name), "rb") as f:
        return f.read()

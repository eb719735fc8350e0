import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library (nvcc cross-compiles without a GPU) and the parity checkers once per session."""
    from bzip2_opencl_b200 import build
    build.build()
    from oracle import oracle, build_ref
    oracle.build()
    build_ref.build()  # no-op when /root/reference is absent (GPU box): the prebuilt oracle/_ref/*.so travel with the repo
    yield


def have_ref(base=100000):
    from oracle import refimpl
    return refimpl.available(base)

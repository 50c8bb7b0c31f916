import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libpoppy_cuda.so; GPU tests fail loudly when it is missing."""
    from poppy_b200 import build, _lib
    build.build()
    return _lib.load()

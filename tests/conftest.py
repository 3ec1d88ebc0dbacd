import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """The C-ABI library; built on demand where the reference checkout exists."""
    from cholespy_b200 import _capi
    if not os.path.exists(_capi.lib_path()):
        from cholespy_b200 import build
        build.build_library()
    return _capi.load()


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle._load()
    return oracle

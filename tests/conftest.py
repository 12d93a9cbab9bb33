import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running trajectory test")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Make sure the C-ABI library and the port oracle exist (builds are seconds)."""
    import __graft_entry__ as ge
    ge.build()

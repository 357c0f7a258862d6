import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def load_package():
    """Import the product package (directory name has hyphens) as icde2019_gpu_join_b200."""
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def gj():
    return load_package()


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure; oracle/fuel_oracle.c)."""
    import oracle
    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def fuel():
    """The product package; GPU tests fail loudly if libfuelgpu.so is missing."""
    import fuel_b200
    fuel_b200.lib()
    return fuel_b200

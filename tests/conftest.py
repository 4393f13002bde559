import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    import nlo_oracle_py
    nlo_oracle_py.build()
    return nlo_oracle_py


@pytest.fixture(scope="session")
def nlo():
    import nonlinear_optimizer_for_slam_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def ctx(nlo):
    c = nlo.Context(0)
    yield c
    c.close()

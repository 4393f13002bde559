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


def pytest_sessionfinish(session, exitstatus):
    """NLO_GUARD=1 python -m pytest tests -m gpu: the whole GPU suite runs with guard bands around every
    device buffer of the library (include/nlo_cuda.h, nlo_debug_guard_report); damaged bands fail the run."""
    if os.environ.get("NLO_GUARD", "0") in ("", "0"):
        return
    pkg = sys.modules.get("nonlinear_optimizer_for_slam_b200")
    if pkg is None:
        return
    rep = pkg.guard_report()
    print("\n[nlo guard] %s" % rep)
    out = os.environ.get("NLO_GUARD_REPORT")
    if out:
        import json
        with open(out, "w") as f:
            json.dump(rep, f)
    if rep["enabled"] and rep["corrupted_bytes"] != 0 and os.environ["NLO_GUARD"] != "selftest":
        session.exitstatus = 1

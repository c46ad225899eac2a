import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """A fresh checkout has no binaries: build the product library, the host layer, the oracle and the CPU build of the
    device math once (no-op when everything is up to date)."""
    need = [os.path.join(ROOT, "auvsl_neural_ode_b200", "libavsim.so"), os.path.join(ROOT, "auvsl_neural_ode_b200", "libavsim_host.so"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so"), os.path.join(ROOT, "tests", "_build", "libhostcheck.so")]
    if not all(os.path.exists(f) for f in need):
        import __graft_entry__
        __graft_entry__.build()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ee():
    """Initialised single-rank eigenexa_b200 (GPU tests only)."""
    import eigenexa_b200 as E
    E.eigen_init(None, "C")
    if E.eigen_get_procs()[0] != 1 or E.last_error():
        raise RuntimeError("eigen_init failed: " + E.last_error())
    yield E
    E.eigen_free()

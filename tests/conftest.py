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
def wmb():
    """The product package (its directory name carries a hyphen, so import it by name)."""
    import importlib
    return importlib.import_module("watermarking-gpu_b200")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle — test infrastructure only (see oracle/wm_oracle.c header)."""
    from oracle import oracle as o
    o.lib()
    return o

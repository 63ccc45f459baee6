import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree once per session (nvcc cross-compiles without a GPU)."""
    from grappa_b200 import build
    if not os.path.exists(build.LIB):
        build.build(verbose=False)
    yield

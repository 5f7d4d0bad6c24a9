import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def libs():
    """The in-tree native libraries; built on demand here (nvcc cross-compiles without a GPU)."""
    from is3d2_b200 import build, capi
    if not os.path.exists(os.path.join(REPO, "is3d2_b200", "libis3d_b200.so")):
        build.build()
    return capi.load_libraries()

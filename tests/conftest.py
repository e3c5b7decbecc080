import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libdxb200.so once per session if it is missing (nvcc cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "diffxpy_b200", "libdxb200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as ge
        ge.build()
    yield

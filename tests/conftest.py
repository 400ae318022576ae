import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle32():
    from oracle import refapi
    return refapi.get("oracle", 32)


@pytest.fixture(scope="session")
def oracle64():
    from oracle import refapi
    return refapi.get("oracle", 64)


@pytest.fixture(scope="session")
def built_libs():
    """The native libraries, built on demand (no-op when the .so files are current)."""
    from opticalflow2d_b200 import build
    return build.build_all()

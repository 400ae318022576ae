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


@pytest.fixture
def exact_engine():
    """For the duration of a test `of.set_strict(False)` selects the EXACT engine (arithmetic level 1: unfused, IEEE
    division, bit-identical fields) instead of the default relaxed one (level 2); afterwards the default is restored."""
    import opticalflow2d_b200 as of
    old = os.environ.get("OF2D_MATH")
    os.environ["OF2D_MATH"] = "exact"
    for bits in (32, 64):
        of.set_math("exact", bits)
    yield
    if old is None:
        os.environ.pop("OF2D_MATH", None)
    else:
        os.environ["OF2D_MATH"] = old
    for bits in (32, 64):
        of.set_strict(False, bits)

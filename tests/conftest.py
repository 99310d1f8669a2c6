import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
# a kernel that never raises its completion flag must end the test run, not hang it (the library itself
# waits without limit unless told otherwise)
os.environ.setdefault("G6_B200_WAIT_SECONDS", "120")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _device_count():
    """CUDA devices the library sees (0 when it is not built or there is no GPU); never opens a device."""
    try:
        from amuse_b200 import g6lib
        return int(g6lib.load().get_device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # g6_open_ exits the process when no CUDA device exists (like sapporo.cpp:24-28): on a CPU-only box a plain
    # `pytest tests` must skip the gpu tests instead of dying in the session fixture
    if _device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def g6():
    """One opened device for the whole GPU test session (the library is a process singleton)."""
    from amuse_b200 import g6lib
    g = g6lib.G6(0)
    yield g
    g.close()

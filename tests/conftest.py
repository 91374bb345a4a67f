import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))                        # checker (tests only)
sys.path.insert(0, os.path.join(ROOT, "gr-liquiddsp_b200", "python"))   # product python layer
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from liquiddsp import capi
        return capi.lib().lqb_device_count() > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def gpu_required():
    if not _have_gpu():
        pytest.fail("a CUDA device and liblqb200.so are required for -m gpu tests (no CPU fallback)")

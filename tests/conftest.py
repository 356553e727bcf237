import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        from bot7_b200 import _lib
        return _lib.lib().b7_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    from bot7_b200 import _lib
    if not _has_gpu():
        pytest.skip("no CUDA device")
    return _lib.Context.default(0)


@pytest.fixture(scope="session")
def oracle():
    import b7_oracle
    return b7_oracle

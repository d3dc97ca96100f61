"""pytest configuration: `gpu` marker, import paths for the product module and the oracle."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "python-liquiddsp_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_ok():
    import liquiddsp
    try:
        return liquiddsp.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda():
    """GPU tests call through the C ABI; without a device they must not silently pass."""
    if not _cuda_ok():
        pytest.fail("no CUDA device: -m gpu tests need the B200 box (there is no CPU fallback to test)")
    return True

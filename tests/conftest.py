import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    import orc as _orc  # oracle bindings: the checker, never the product
    return _orc


@pytest.fixture(scope="session")
def rt(orc):
    return orc.rt


@pytest.fixture(scope="session")
def gpu(rt):
    """The product library on a real device; fails loudly instead of falling back."""
    L = rt.product_lib()
    if L.rt_device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box (there is no CPU fallback)")
    return L

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_cuda_device():
    try:
        from specfab_b200 import _lib
        return _lib.load().sfb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing them with SFB_ECUDA
    (the product has no CPU fallback, so they cannot run there); `-m gpu` on the B200 box runs them all."""
    if _have_cuda_device():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the library has no CPU fallback")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)

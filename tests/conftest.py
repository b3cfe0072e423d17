import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import ctypes
        n = ctypes.c_int(0)
        return ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a CUDA device: the gpu-marked tests are skipped, not failed"""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def golden():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def scene_json():
    def get(name):
        return os.path.join(ROOT, "scenes", "json", name + ".json")
    return get

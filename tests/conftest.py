import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _simt_check():
    """True inside the subprocess that tests/test_simt_check.py starts: the gpu-marked parity tests then
    run against the kernels interpreted on the CPU (tests/simt), which needs no device."""
    return os.environ.get("CPIC_B200_SIMT_CHECK") == "1" and "simt" in os.environ.get("CPIC_B200_LIB", "")


def pytest_collection_modifyitems(config, items):
    if _have_gpu() or _simt_check():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def root():
    return ROOT


def conf_path(name):
    return os.path.join(ROOT, "conf", name)

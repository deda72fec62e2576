import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (runs on the B200 box)")


def _cuda_device_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        import shutil
        import subprocess
        if not shutil.which("nvidia-smi"):
            return False
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not failed) on a box without a CUDA device."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the library has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")

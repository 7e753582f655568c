import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        from decipher_b200.capi import load_library
        return load_library().dcp_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests must FAIL loudly on a GPU box whose extension is missing; on a CPU-only box they are
    # deselected by `-m "not gpu"`.  If someone runs them without a device, skip with the reason.
    if _have_gpu():
        return
    import shutil
    if shutil.which("nvidia-smi") and os.path.exists("/dev/nvidia0"):
        return      # a GPU exists but the library did not load: let the tests fail
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _cuda_ok():
    """ask the product library itself (it is ctypes-only and does not need torch): pgpu_create succeeds iff a CUDA
    device is usable"""
    try:
        from pyrodigal_b200 import _capi

        _capi.Context(0).close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if not any("gpu" in item.keywords for item in items) or _cuda_ok():
        return
    if (config.getoption("-m") or "").strip() == "gpu" and os.environ.get("PGPU_ALLOW_GPU_SKIP") != "1":
        # an explicit GPU run without a usable device must not read as "passed"
        raise pytest.UsageError("-m gpu: pyrodigal_b200 could not create a context on CUDA device 0 "
                                "(library missing or no GPU); set PGPU_ALLOW_GPU_SKIP=1 to skip instead")
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        from minivectordb_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # No CUDA device: GPU tests are SKIPPED (the build container has none), unless the caller asked
    # for certainty with MVDB_REQUIRE_GPU=1 -- then a GPU-less run of GPU tests is an error, so that a
    # CI job that is supposed to run on a B200 cannot go green with every parity test skipped.
    if _has_gpu():
        return
    if os.environ.get("MVDB_REQUIRE_GPU") == "1" and any("gpu" in item.keywords for item in items):
        raise pytest.UsageError("MVDB_REQUIRE_GPU=1 but no CUDA device is visible: GPU tests cannot run")
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from svdss_b200 import capi
        return capi.lib().svb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # gpu-marked tests fail loudly (not skip) on a box with a GPU whose library is missing; on a
    # box without any GPU they are skipped so that a plain `pytest tests/` stays usable.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)

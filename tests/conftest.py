import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _device_count() -> int:
    # a missing libpik_b200.so must fail loudly, not skip: no try/except here
    from pick_ik_b200 import capi

    return capi.device_count()


def pytest_collection_modifyitems(config, items):
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    if _device_count() == 0:
        skip = pytest.mark.skip(reason="no CUDA device in this container")
        for it in gpu_items:
            it.add_marker(skip)

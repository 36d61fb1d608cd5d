import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def lib_option():
    """Set path-selection switches of libb200_roi_ops.so for one test (b200_roi_ops_set_option); restored to the default
    afterwards.  The library reads the environment only once, at first use, so tests cannot steer it with setenv."""
    from detectron.pytorch_b200 import _lib
    touched = []

    def set_option(name, value):
        _lib.set_option(name, value)
        touched.append(name)

    yield set_option
    for name in touched:
        _lib.set_option(name, None)

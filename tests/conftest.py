import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line(
        "markers", "reference: test imports /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle.ref_loader import reference_available
    if reference_available():
        return
    skip = pytest.mark.skip(reason="reference checkout not present")
    for item in items:
        if "reference" in item.keywords:
            item.add_marker(skip)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# No test here needs more than a couple of minutes on the B200 box (the whole GPU suite is ~2.5 min).
# One run on a box did not come back from the 4th test in 10 minutes (profiles/README.md): with
# pytest-timeout present a stuck test now dumps every thread's stack and ends the run instead.
TEST_TIMEOUT_S = 420


def pytest_collection_modifyitems(config, items):
    if config.pluginmanager.hasplugin("timeout"):
        for item in items:
            if item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(TEST_TIMEOUT_S, method="thread"))
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

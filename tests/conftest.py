import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the reference's own code (np.matrix, np.in1d) and the np.matrix the drop-in API returns, like it
    config.addinivalue_line("filterwarnings", "ignore:the matrix subclass:PendingDeprecationWarning")
    config.addinivalue_line("filterwarnings", "ignore:`in1d` is deprecated:DeprecationWarning")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

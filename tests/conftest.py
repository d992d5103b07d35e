import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'ref: needs oracle/_ref (the reference C++ built from /root/reference)')


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    gpu = _have_gpu()
    from oracle import reflib
    have_ref = reflib.available()
    for item in items:
        if 'gpu' in item.keywords and not gpu:
            item.add_marker(pytest.mark.skip(reason='no CUDA device'))
        if 'ref' in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason='oracle/_ref not built (needs /root/reference)'))

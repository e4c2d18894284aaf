import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    has_ref = os.path.isdir("/root/reference/src")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    has_lib = os.path.exists(os.path.join(ROOT, "multivae_b200", "libmultivae_b200.so"))
    skip_gpu = pytest.mark.skip(reason="needs a CUDA device and the built libmultivae_b200.so (run on the B200 box with -m gpu)")
    for item in items:
        if "reference" in item.keywords and not has_ref:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords and not (has_gpu and has_lib):
            item.add_marker(skip_gpu)

import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import refenv  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line(
        "markers", "reference: needs the real reference (baseline/_ref, written by "
        "tools/vendor_reference.py, or /root/reference in the authoring container)"
    )


def _cuda_ready():
    try:
        import torch

        return torch.cuda.is_available() and os.path.exists(
            os.path.join(REPO, "torchtree_b200", "lib", "libttb200.so"))
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have_ref = refenv.available()
    skip_ref = pytest.mark.skip(reason="the reference is not available (tools/vendor_reference.py)")
    have_gpu = _cuda_ready()
    skip_gpu = pytest.mark.skip(reason="needs a CUDA device and the built libttb200.so")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(skip_gpu)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _have_gpu():
    # tests/test_emu_parity.py re-runs a subset of the `gpu` tests in a subprocess against the CPU
    # lock-step emulation build of the same sources (tests/cuemu): test infrastructure only
    if os.environ.get("BENDY_CUDA_EMU") == "1" and os.environ.get("BENDY2D_B200_LIB"):
        return True
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build the product library (nvcc, sm_100a) and the C oracle once per session."""
    from galahad_b200 import build as B
    B.build()
    from oracle import refssids
    refssids.build_oracle()
    if not refssids.have_ref() and os.path.isdir("/root/reference/src/ssids/cpu"):
        refssids.build_ref()
    return True


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False

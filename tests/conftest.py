import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One frb_ctx for the whole GPU session (fp16 activations, the default)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from facerec_b200._lib import Context
    c = Context(0)
    yield c
    c.close()

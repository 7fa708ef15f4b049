import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_lib():
    """Initialise the CUDA library once per session on cuda:0 (single rank)."""
    from svfsi_b200 import api
    api.init(device=0, rank=0, nranks=1)
    yield api
    api.finalize()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.lib()  # builds liblkoracle.so on first use
    return O


@pytest.fixture(scope="session")
def ctx():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cuda_optical_flow_2_b200 import Context

    c = Context(0)
    yield c
    c.close()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import spotv2net_b200
    return spotv2net_b200.load_library()      # raises (does not skip) if the .so is missing

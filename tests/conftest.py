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


def pytest_terminal_summary(terminalreporter):
    """Report every (test, tensor) that passed on the 3x-fp32-oracle allowance of tests/test_gpu_parity.py."""
    mod = sys.modules.get("test_gpu_parity") or sys.modules.get("tests.test_gpu_parity")
    used = getattr(mod, "ALLOWANCE_USED", None) if mod else None
    if used:
        terminalreporter.write_line(f"parity allowance (3x fp32 oracle) used by {len(used)} (test, tensor) pairs:")
        for t, k, e, e32 in used:
            terminalreporter.write_line(f"  {t}  {k}  ours {e:.2e}  fp32 oracle {e32:.2e}")

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `pytest -m gpu`")


@pytest.fixture(scope="session")
def cuda_lib():
    """The C-ABI library, on a GPU box.  Fails loudly (no fallback) if it is missing or the device is not sm_100."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test on a box without CUDA"
    from sdvar_b200 import _cabi
    l = _cabi.lib()
    rc = l.sdvar_arch_check(0)
    assert rc == 0, l.sdvar_last_error().decode()
    return _cabi

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def sq():
    """The product package with its CUDA library loaded on cuda:0 (gpu tests only)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sequitr_b200
    sequitr_b200.require_gpu()
    return sequitr_b200

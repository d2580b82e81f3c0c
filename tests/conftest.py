import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REFERENCE = '/root/reference'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA B200 device (run on the GPU box with -m gpu)')


def has_reference():
    return os.path.isdir(os.path.join(REFERENCE, 'dust3r'))


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')

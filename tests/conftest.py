import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def cuda_lib():
    """The built C-ABI library; GPU tests must fail (not skip) when it is missing."""
    import torch
    assert torch.cuda.is_available(), 'GPU test selected but no CUDA device'
    from sgcdet_b200 import _lib, build
    build.build()   # no-op when the in-tree library is up to date (it normally travels with the snapshot)
    return _lib.load()

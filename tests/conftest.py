import os
import sys

import pytest

# a peer-memory wait that can never be satisfied (a bug) should cost a test seconds, not the library's 30 s default
os.environ.setdefault('KT_P2P_TIMEOUT_MS', '10000')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def lib_path():
    """Build (if stale) and return libkarman_b200.so -- nvcc cross-compiles without a GPU."""
    from karman_taichi_b200 import build
    return build.build()


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from karman_taichi_b200 import build
    build.build()
    return torch.device('cuda', 0)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def gpu():
    """Bind the process to cuda:0 through the C ABI; fail loudly if that is impossible."""
    from dynamite_b200 import _capi
    _capi.ensure_gpu(0)
    return _capi

import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def built_lib():
    """Path of libtss_b200.so, building it (nvcc cross-compiles without a GPU) if absent."""
    from torch_semantic_segmentation_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(['make', '-j8'], cwd=ROOT)
    return _lib.LIB_PATH


@pytest.fixture
def fake_backend():
    """Install the CPU emulation of the C ABI (host-logic tests only)."""
    from torch_semantic_segmentation_b200 import _lib
    from tests.fake_backend import FakeBackend
    prev = _lib._backend
    fb = FakeBackend()
    _lib.set_backend(fb)
    yield fb
    _lib.set_backend(prev)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (PKG, os.path.join(ROOT, 'oracle'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: no CPU fallback exists
    pass


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN

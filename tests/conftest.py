import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'ms-eetc_b200')
for p in (ROOT, PKG, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'hostsim')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def data_dir():
    return PKG


@pytest.fixture(scope='session')
def built_lib():
    "Compile libmseetc_b200.so if it is stale (nvcc cross-compiles without a GPU)."
    import __graft_entry__ as ge
    ge.build()
    return ge.LIB

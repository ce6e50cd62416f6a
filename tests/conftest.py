import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'kapre_cases.npz'))


def pytest_sessionstart(session):
    """A fresh checkout has no built library (it is git-ignored): build it once if nvcc is here, so that the
    ABI / emulator tests do not depend on someone having called __graft_entry__.build() first."""
    import shutil
    lib = os.path.join(ROOT, 'kapre_b200', '_lib', 'libkapre_b200.so')
    if not os.path.exists(lib) and shutil.which(os.environ.get('NVCC', 'nvcc')):
        import __graft_entry__
        __graft_entry__.build()

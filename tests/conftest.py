import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope='session')
def ethanol():
    g = golden('ethanol_32.npz')
    return {'R': g['R'], 'z': g['z']}


@pytest.fixture(scope='session')
def solid():
    g = golden('solid_0.npz')
    return {'R': g['R'], 'z': g['z'], 'unit_cell': g['unit_cell']}


@pytest.fixture(scope='session')
def emu_lib_path():
    """CPU-emulated build of the SIMT kernels (test infrastructure, see tests/emu/cuda_emu.h)."""
    from emu_engine import build_emu
    return build_emu()


@pytest.fixture(scope='session')
def cuda_lib_path():
    """The product library; cross-compiled here with nvcc (no GPU needed to build or dlopen it)."""
    from mlff_b200 import _lib
    return _lib.build()

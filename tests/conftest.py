import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_case(name):
    """Golden case written by tests/golden/make_golden.py -> (arrays, weights state dict)."""
    d = dict(np.load(os.path.join(GOLDEN, f'case_{name}.npz')))
    w = dict(np.load(os.path.join(GOLDEN, f'weights_{d["weights"]}.npz')))
    return d, w


def load_weights(name):
    return dict(np.load(os.path.join(GOLDEN, f'weights_{name}.npz')))


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN

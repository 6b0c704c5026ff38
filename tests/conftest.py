import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    class G:
        def npz(self, name):
            return np.load(os.path.join(GOLD, name + '.npz'))

        def json(self, name):
            with open(os.path.join(GOLD, name + '.json')) as f:
                return json.load(f)
    return G()


def max_rel(a, b):
    """max|a-b| / max|b| -- the parity metric SURVEY.md §8d defines."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

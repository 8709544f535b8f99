import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')
    config.addinivalue_line('markers', 'slow: full-size cases')


def golden_path(name):
    return os.path.join(ROOT, 'tests', 'golden', f'{name}.npz')


@pytest.fixture(scope='session')
def state_dicts():
    """Seeded hot-path weights, cached per num_layers."""
    from mv2d_b200 import synth
    cache = {}

    def get(num_layers=6):
        if num_layers not in cache:
            cache[num_layers] = synth.make_state_dict(0, num_layers=num_layers)
        return cache[num_layers]
    return get

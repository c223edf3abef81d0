import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu under gpurun)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope='session')
def real_weights():
    """The shipped RAFT-OU checkpoint through the oracle's loader (skip if it did not travel)."""
    from oracle import fetch_ref_assets, mft_oracle
    path = fetch_ref_assets.find_checkpoint()
    if path is None:
        pytest.skip('shipped checkpoint not available on this box')
    return mft_oracle.load_checkpoint(path)


@pytest.fixture(scope='session')
def seeded_weights():
    from oracle import mft_oracle
    return mft_oracle.seeded_weights(0)

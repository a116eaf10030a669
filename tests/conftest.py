"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` tests: oracle vs. the committed golden vectors (generated from the reference's own
Python, tests/golden/make_golden.py), host logic, C-ABI symbol export.
`-m gpu` tests: parity of the CUDA path (through the C-ABI) against the oracle and the goldens.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def fixture_spec():
    return json.load(open(os.path.join(GOLDEN_DIR, 'fixture_main_map.json')))


@pytest.fixture(scope='session')
def golden():
    return np.load(os.path.join(GOLDEN_DIR, 'golden_ref.npz'))


@pytest.fixture(scope='session')
def golden_meta():
    return json.load(open(os.path.join(GOLDEN_DIR, 'golden_ref_meta.json')))


def full_paths(spec, X):
    """[x_start, x, x_goal] per row -> the reference's z_ layout (solver.py:64-66)."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    s = np.broadcast_to(np.asarray(spec['x_start'], dtype=np.float64), (X.shape[0], 2))
    g = np.broadcast_to(np.asarray(spec['x_goal'], dtype=np.float64), (X.shape[0], 2))
    return np.concatenate([s, X, g], axis=1)

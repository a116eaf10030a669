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


def build_product_map(spec):
    """fixture spec -> the product's RegionMap + Problem (through the reference-named constructors)."""
    import uam_path_planning_b200 as uam
    mk = {'polygon': lambda s: uam.polygon(*s['verts']),
          'ball': lambda s: uam.ball(s['center'], s.get('r1'), s.get('r2')),
          'square': lambda s: uam.square(s['center'], s['r1'], s.get('r2'))}
    m = uam.RegionMap()
    m.add_obstacles(*[mk[s['kind']](s) for s in spec['obstacles']])
    for name, shapes in spec['regions']:
        m.new_region(name, 'red')
        m.add_shapes_to_region(name, *[mk[s['kind']](s) for s in shapes])
    m.x_start, m.x_goal = spec['x_start'], spec['x_goal']
    return m


def build_product_problem(spec, N, options=None, weights=None, enlargement=None):
    import uam_path_planning_b200 as uam
    m = build_product_map(spec)
    prob = uam.Problem(m, N, dict(spec['options'], **(options or {})))
    prob.params.update(maxratio=spec['maxratio'], maxalpha=spec['maxalpha'],
                       enlargement=spec['enlargement'] if enlargement is None else enlargement)
    for (name, _), w in zip(spec['regions'], weights or spec['weights']):
        prob.set_weight(name, w)
    return prob


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False

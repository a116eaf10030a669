"""INTEGRATION.md section 2 executed: integration/gpu_backend.py (the file a maintainer adds to the reference) binds
libuam_b200.so with plain ctypes from the reference's own classes.

CPU part (needs /root/reference, skipped elsewhere): the tables read out of the reference's closures equal the frozen
tests/golden/reference_tables.npz and, bit for bit, the tables the product's own constructors build.
GPU part: the raw binding, fed with the frozen tables, reproduces the reference's golden costs / constraint vectors.
"""
import os
import sys

import numpy as np
import pytest

from conftest import build_product_map, full_paths

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'integration'))
GOLD = os.path.join(ROOT, 'tests', 'golden')
HAVE_REF = os.path.isdir('/root/reference/geo_simulation_project/path_generation')


def _tables():
    t = np.load(os.path.join(GOLD, 'reference_tables.npz'))
    return {k: t[k] for k in t.files}


@pytest.mark.skipif(not HAVE_REF, reason='needs the reference tree (authoring container only)')
def test_tables_from_reference_closures(fixture_spec):
    import subprocess
    # in a subprocess: the reference's modules and the casadi stand-in must not leak into this interpreter
    code = ('import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r);'
            'import make_reference_tables as mt, gpu_backend as gb;'
            'spec, m, pr = mt.reference_map_and_problem();'
            't = gb.extract_tables(m); g = gb.GpuProblem.__new__(gb.GpuProblem); g.p = pr;'
            'np.savez(sys.argv[1], p=g.parameter_vector(), flags=g.flags(), **t)') % (GOLD, os.path.join(ROOT, 'integration'))
    import tempfile
    out = os.path.join(tempfile.mkdtemp(), 't.npz')
    subprocess.run([sys.executable, '-W', 'ignore', '-c', code, out], check=True, cwd=GOLD)
    got, frozen = np.load(out), _tables()
    for k in ('edges', 'off', 'region', 'center'):
        assert np.array_equal(got[k], frozen[k]), k
    assert int(got['n_regions']) == int(frozen['n_regions']) == 3
    # ... and they are the tables the product's constructors build from the same numbers
    from uam_path_planning_b200.shapes import flatten_shapes
    m = build_product_map(fixture_spec)
    edges, off, reg, cen = flatten_shapes(m.obstacles, m._region_lists())
    assert np.array_equal(edges, frozen['edges']) and np.array_equal(off, frozen['off'])
    assert np.array_equal(reg, frozen['region']) and np.array_equal(cen, frozen['center'])
    # parameter vector / flags in the reference's order (solver.py:60-68, problem.py:12-17)
    f = fixture_spec
    assert np.array_equal(got['p'], np.array([*f['x_start'], *f['x_goal'], f['maxratio'], f['maxalpha'], f['enlargement'], *f['weights']]))
    assert int(got['flags']) == 0b0111


def test_frozen_tables_match_product_tables(fixture_spec):
    """runs everywhere: the frozen reference tables == the product's tables (incl. the box-side records of square())"""
    import uam_path_planning_b200 as uam
    from uam_path_planning_b200.shapes import flatten_shapes
    t = _tables()
    m = build_product_map(fixture_spec)
    edges, off, reg, cen = flatten_shapes(m.obstacles, m._region_lists())
    assert np.array_equal(edges, t['edges']) and np.array_equal(off, t['off'])
    assert np.array_equal(reg, t['region']) and np.array_equal(cen, t['center'])
    assert np.array_equal(uam.square([1.0, 1.0], 0.5, 0.25).records(), t['square_records'])


@pytest.mark.gpu
@pytest.mark.parametrize('N', [80, 62, 5])
def test_raw_ctypes_binding_matches_reference_goldens(fixture_spec, golden, N):
    import gpu_backend as gb
    from uam_path_planning_b200 import build
    f = fixture_spec
    sc = gb.GpuScorer(_tables(), lib=gb.load_library(build.OUT))
    p = np.array([*f['x_start'], *f['x_goal'], f['maxratio'], f['maxalpha'], f['enlargement'], *f['weights']])
    Z = full_paths(f, golden[f'arc_N{N}_x'])
    cost, col, g = sc.score(Z, N, p, 0b0111, want_g=True)
    np.testing.assert_allclose(cost, golden[f'arc_N{N}_cost'], rtol=1e-12)
    assert np.array_equal(col, golden[f'arc_N{N}_collide'].any(axis=1))
    np.testing.assert_allclose(g, golden[f'arc_N{N}_g'], rtol=1e-9, atol=1e-300)
    assert np.array_equal(g == 0, golden[f'arc_N{N}_g'] == 0)
    sc.close()

#!/usr/bin/env python
"""Regenerate tests/golden/* by running the REFERENCE'S OWN Python on this machine.

Not framework code.  Run only in the authoring container, where /root/reference exists:

    python tests/golden/make_golden.py

The reference's path_generation modules import `casadi`, `matplotlib` and `opengen`, none of which
is installed here, and only use casadi's fmin/fmax/norm_2/sumsqr/dot/cos/sqrt/vertcat/reshape on
plain floats for the numeric path.  This script writes a numpy stand-in for those names into a
temporary directory, puts it on sys.path in front of the reference's sources, imports the
reference's problem.py / region_map.py / map.py / quadratic_obstacle.py / polygon.py / ball.py /
square.py / solver.py unmodified and records what they return.  All arithmetic
therefore runs through the reference's code, in its operation order, in IEEE double.

Outputs (committed):
    tests/golden/fixture_main_map.json   map spec = numbers of path_generation/main.py:27-49,128-145
                                         and data/processed/{land,populated}_area.txt
    tests/golden/golden_ref.npz          reference outputs on that fixture
    tests/golden/golden_ref_meta.json    scalar goldens and provenance
"""
import json
import os
import re
import sys
import tempfile
import textwrap

import numpy as np

REF = '/root/reference'
PG = os.path.join(REF, 'geo_simulation_project', 'path_generation')
HERE = os.path.dirname(os.path.abspath(__file__))

CASADI_SHIM = textwrap.dedent('''
    import numpy as _np
    def fmin(a, b): return _np.minimum(a, b)
    def fmax(a, b): return _np.maximum(a, b)
    def norm_2(a): return _np.linalg.norm(_np.asarray(a, dtype=float).ravel())
    def sumsqr(a):
        a = _np.asarray(a, dtype=float).ravel()
        return _np.sum(a * a)
    def dot(a, b): return float(_np.dot(_np.asarray(a, dtype=float).ravel(), _np.asarray(b, dtype=float).ravel()))
    def cos(a): return _np.cos(a)
    def sqrt(a): return _np.sqrt(a)
    def vertcat(*a): return _np.concatenate([_np.asarray(x, dtype=float).ravel() for x in a]) if a else _np.zeros(0)
    def reshape(a, shape): return _np.asarray(a, dtype=float).ravel()
    def DM(a): return _np.asarray(a, dtype=float)
    class SX:  pass
    class MX:  pass
''')

MPL_PYPLOT = textwrap.dedent('''
    class _Any:
        def __getattr__(self, k): return _Any()
        def __call__(self, *a, **k): return _Any()
    def __getattr__(name): return _Any()
''')
MPL_PATCHES = "class Polygon: pass\nclass Circle: pass\nclass Patch: pass\n"


def build_shim(d):
    os.makedirs(os.path.join(d, 'casadi'))
    open(os.path.join(d, 'casadi', '__init__.py'), 'w').write(CASADI_SHIM)
    open(os.path.join(d, 'casadi', 'casadi.py'), 'w').write(CASADI_SHIM)
    os.makedirs(os.path.join(d, 'matplotlib'))
    open(os.path.join(d, 'matplotlib', '__init__.py'), 'w').write('')
    open(os.path.join(d, 'matplotlib', 'pyplot.py'), 'w').write(MPL_PYPLOT)
    open(os.path.join(d, 'matplotlib', 'patches.py'), 'w').write(MPL_PATCHES)
    os.makedirs(os.path.join(d, 'opengen'))
    open(os.path.join(d, 'opengen', '__init__.py'), 'w').write('')


def parse_polygons(path):
    """numbers only: 'vertices = [polygon([x, y], ...), ...]' -> list of vertex lists"""
    txt = open(path).read()
    out = []
    for body in re.findall(r'polygon\((.*?)\)', txt, flags=re.S):
        out.append([[float(a), float(b)] for a, b in re.findall(r'\[\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*\]', body)])
    return out


def main():
    shim = tempfile.mkdtemp(prefix='uam_shim_')
    build_shim(shim)
    sys.path.insert(0, PG)
    sys.path.insert(0, shim)
    os.chdir(PG)

    from region_map import RegionMap
    from problem import Problem
    from solver import Solver
    from polygon import polygon
    from ball import ball
    from square import square
    import utils as ut

    data = os.path.join(REF, 'data', 'processed')
    # ---- map of path_generation/main.py:21-49 ------------------------------------------------
    discs = [([38.66652661075855, -9.203164091309498], 9), ([46.36137256675563, 3.9427562315386298], 2),
             ([19.846825121034392, 18.93411773399299], 2), ([26.037433469490207, 15.46710452712196], 2),
             ([46.87758543585609, -19.138710035318375], 2)]
    m = RegionMap()
    m.add_obstacles(*[ball(c, r) for c, r in discs])
    m.new_region('Land', [0.9290, 0.6940, 0.1250])
    m.add_shapes_to_region('Land', *ut.get_var_from_file(os.path.join(data, 'land_area.txt'), 'vertices'))
    m.new_region('Population', 'Red')
    m.add_shapes_to_region('Population', *ut.get_var_from_file(os.path.join(data, 'populated_area.txt'), 'vertices'))
    m.new_region('HistCenter', 'Green')
    m.add_shape_to_region('HistCenter', ball([33.874752, -24.981154], 1))
    x_start, x_goal = [35.590685, -27.711422], [26.478673, 9.564082]
    m.x_start, m.x_goal = x_start, x_goal
    weights = [200, 15000, 27000]
    maxratio, maxalpha, enlargement = 1.04, np.pi / 80, 0
    opts = {'length_smooth': True, 'penalty_smooth': True, 'obstacle_smooth': True, 'maxratio_smooth': False}

    fixture = {
        'source': 'path_generation/main.py:27-49,128-145 + data/processed/{land,populated}_area.txt',
        'obstacles': [{'kind': 'ball', 'center': c, 'r1': r, 'r2': r} for c, r in discs],
        'regions': [
            ['Land', [{'kind': 'polygon', 'verts': v} for v in parse_polygons(os.path.join(data, 'land_area.txt'))]],
            ['Population', [{'kind': 'polygon', 'verts': v} for v in parse_polygons(os.path.join(data, 'populated_area.txt'))]],
            ['HistCenter', [{'kind': 'ball', 'center': [33.874752, -24.981154], 'r1': 1, 'r2': 1}]],
        ],
        'x_start': x_start, 'x_goal': x_goal, 'weights': weights,
        'maxratio': maxratio, 'maxalpha': maxalpha, 'enlargement': enlargement, 'options': opts,
        'no_fly_area_txt': parse_polygons(os.path.join(data, 'no_fly_area.txt')),
    }
    json.dump(fixture, open(os.path.join(HERE, 'fixture_main_map.json'), 'w'), indent=1)

    def mk_problem(N, o=None, e=enlargement, w=weights):
        pr = Problem(m, N, dict(opts, **(o or {})))
        pr.params.update({'maxratio': maxratio, 'maxalpha': maxalpha, 'enlargement': e})
        for name, wi in zip(m.region_names(), w):
            pr.set_weight(name, wi)
        return pr

    def full(x):
        return np.concatenate([x_start, x, x_goal])

    out = {}
    meta = {'generated_by': 'tests/golden/make_golden.py', 'reference': 'nomaporon/uam_path_planning @ /root/reference',
            'numpy': np.__version__}

    # ---- arcs (B.1) ---------------------------------------------------------------------------
    disp = [-0.5, -0.25, 0.0, 0.25, 0.5]
    for N in (80, 62, 64, 5):
        pr = mk_problem(N)
        sv = Solver(pr, {})
        X = np.stack([sv.create_x_init(d) for d in disp])
        out[f'arc_N{N}_x'] = X
        out[f'arc_N{N}_cost'] = np.array([float(pr.get_cost(full(x))) for x in X])
        out[f'arc_N{N}_g'] = np.stack([np.asarray(pr.get_nonlincon(full(x)), dtype=float) for x in X])
        out[f'arc_N{N}_length'] = np.array([float(pr.length_of(x)) for x in X])
        out[f'arc_N{N}_length_smooth'] = np.array([float(pr.length_of(x, True)) for x in X])
        out[f'arc_N{N}_collide'] = np.array([[bool(m.collides(full(x)[2 * j:2 * j + 2])) for j in range(N + 2)] for x in X])
        print('arcs N', N, out[f'arc_N{N}_cost'])
    out['arc_disp'] = np.array(disp)

    # ---- jittered paths, N = 62 (W = 64), the C2/C3 path shape ----------------------------------
    N = 62
    pr = mk_problem(N)
    sv = Solver(pr, {})
    rng = np.random.default_rng(123)
    J = []
    for i in range(24):
        d = float(rng.uniform(-0.9, 0.9))
        J.append(sv.create_x_init(d) + rng.normal(0, 0.05 + 0.3 * (i % 3), 2 * N))
    J = np.stack(J)
    out['jit_x'] = J
    out['jit_cost'] = np.array([float(pr.get_cost(full(x))) for x in J])
    out['jit_g'] = np.stack([np.asarray(pr.get_nonlincon(full(x)), dtype=float) for x in J])
    out['jit_collide'] = np.array([[bool(m.collides(full(x)[2 * j:2 * j + 2])) for j in range(N + 2)] for x in J])
    # SURVEY B.1 jitter case
    rng = np.random.default_rng(123)
    xj = sv.create_x_init(0) + rng.normal(0, 0.05, 2 * N)
    meta['survey_jitter_cost'] = float(pr.get_cost(full(xj)))
    out['survey_jitter_x'] = xj
    out['survey_jitter_g'] = np.asarray(pr.get_nonlincon(full(xj)), dtype=float)

    # ---- option / parameter variants, straight line N = 80 --------------------------------------
    N = 80
    x0 = Solver(mk_problem(N), {}).create_x_init(0)
    var = {}
    var['enlargement1'] = float(mk_problem(N, e=1).get_cost(full(x0)))
    var['length_nonsmooth'] = float(mk_problem(N, {'length_smooth': False}).get_cost(full(x0)))
    with np.errstate(all='ignore'):
        var['penalty_nonsmooth'] = float(np.asarray(mk_problem(N, {'penalty_smooth': False}).get_cost(full(x0))).ravel()[0])
    var['weights_alt'] = float(mk_problem(N, w=[100, 7500, 13500]).get_cost(full(x0)))
    var['enlargement_neg'] = float(mk_problem(N, e=-0.25).get_cost(full(x0)))
    meta['variants_straight_N80'] = var
    xa = Solver(mk_problem(N), {}).create_x_init(0.25)
    out['var_g_obstacle_nonsmooth'] = np.asarray(mk_problem(N, {'obstacle_smooth': False}).get_nonlincon(full(xa)), dtype=float)
    out['var_g_maxratio_smooth'] = np.asarray(mk_problem(N, {'maxratio_smooth': True}).get_nonlincon(full(xa)), dtype=float)
    out['var_x'] = xa

    # ---- point queries (B.2) --------------------------------------------------------------------
    pr = mk_problem(80)
    rng = np.random.default_rng(7)
    Q = np.concatenate([
        np.array([[33.874752, -24.981154], [35.590685, -27.711422], [30, -20],
                  [38.66652661075855, -9.203164091309498], [0, 0], [33, -25], [40, -9]], dtype=float),
        np.column_stack([rng.uniform(8, 72, 200), rng.uniform(-42, 22, 200)]),
        np.array(discs[1][0]) + 2.0 * np.column_stack([np.cos(np.linspace(0, 6.28, 40)), np.sin(np.linspace(0, 6.28, 40))])
        * (1 + rng.normal(0, 1e-3, (40, 1))),
    ])
    out['pt_x'] = Q
    tp = pr.get_total_penalty_function()
    out['pt_total_penalty'] = np.array([float(np.asarray(tp(q)).ravel()[0]) for q in Q])
    out['pt_region_penalty'] = np.array([[float(np.asarray(pr.get_penalty_function(r)(q)).ravel()[0]) for r in m.region_names()] for q in Q])
    out['pt_obstacle_penalty'] = np.array([float(np.asarray(pr.get_penalty_function(None)(q)).ravel()[0]) for q in Q])
    out['pt_collides'] = np.array([bool(m.collides(q)) for q in Q])
    out['pt_getitem'] = np.array([bool(m[(float(q[0]), float(q[1]))]) for q in Q])

    # ---- raw inequality values: pins edge order, sign and arithmetic of every shape --------------
    shapes = list(m.obstacles) + [s for r in m.region_names() for s in m.regions[r]['shapes']]
    Hq = Q[:64]
    hv, hoff = [], [0]
    for s in shapes:
        for h in s.inequalities:
            hv.append([float(np.asarray(h(q)).ravel()[0]) for q in Hq])
        hoff.append(len(hv))
    out['h_values'] = np.array(hv)
    out['h_offsets'] = np.array(hoff)
    out['h_points'] = Hq
    out['shape_centers'] = np.array([np.asarray(s.center, dtype=float).ravel() for s in shapes])
    out['shape_areas'] = np.array([float(s.area) for s in shapes])
    out['shape_psi_center'] = np.array([float(np.asarray(s.penalty_function(True, 0)(s.center)).ravel()[0]) for s in shapes])

    # ---- constructors: square / ball / polygon behaviour (B.2) ------------------------------------
    sq = square([1, 1], 0.5)
    bl = ball([1, 1], 2, 1)
    cons = {
        'square_contains_1.2_0.9': bool(sq.contains(np.array([1.2, 0.9]))),
        'square_contains_1.6_1': bool(sq.contains(np.array([1.6, 1.0]))),
        'square_psi_1.2_0.9': float(sq.penalty_function(True, 0)(np.array([1.2, 0.9]))),
        'ball_contains_2.9_1': bool(bl.contains(np.array([2.9, 1.0]))),
        'ball_contains_1_2.1': bool(bl.contains(np.array([1.0, 2.1]))),
        'ball_area': float(bl.area), 'square_area': float(sq.area),
        'unit_square_edge_1e-15': bool(polygon([0., 0.], [1., 0.], [1., 1.], [0., 1.]).contains(np.array([1 + 1e-15, .5]))),
        'unit_square_edge_1e-13': bool(polygon([0., 0.], [1., 0.], [1., 1.], [0., 1.]).contains(np.array([1 + 1e-13, .5]))),
    }
    errs = {}
    for name, pts in {'two_vertices': [[0., 0.], [1., 1.]], 'aligned': [[0., 0.], [1., 0.], [2., 0.], [1., 1.]],
                      'nonconvex': [[0., 0.], [2., 0.], [0.5, 0.5], [0., 2.]]}.items():
        try:
            polygon(*pts)
            errs[name] = None
        except Exception as ex:  # noqa
            errs[name] = [type(ex).__name__, str(ex)]
    try:
        Solver(mk_problem(10), {}).create_x_init(1.5)
        errs['x_init_1.5'] = None
    except Exception as ex:  # noqa
        errs['x_init_1.5'] = [type(ex).__name__, str(ex)]
    cons['errors'] = errs
    meta['constructors'] = cons
    # polygon with shuffled vertex order: pins the gift-wrap
    pv = [[28.836, -32.708], [32.53, -35.464], [29.607, -36.124], [31.759, -32.048]]
    pg = polygon(*pv)
    out['shuffled_poly_verts'] = np.array(pv)
    out['shuffled_poly_h'] = np.array([[float(np.asarray(h(q)).ravel()[0]) for q in Hq] for h in pg.inequalities])
    out['shuffled_poly_center'] = np.asarray(pg.center, dtype=float).ravel()
    meta['shuffled_poly_area'] = float(pg.area)

    np.savez_compressed(os.path.join(HERE, 'golden_ref.npz'), **out)
    json.dump(meta, open(os.path.join(HERE, 'golden_ref_meta.json'), 'w'), indent=1)
    print(json.dumps(meta, indent=1))


if __name__ == '__main__':
    main()

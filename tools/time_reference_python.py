#!/usr/bin/env python
"""Time the REFERENCE'S OWN Python on its own scenario (authoring container only: needs /root/reference).

    python tools/time_reference_python.py [n_paths]

Imports the unmodified path_generation modules under the numpy stand-in for CasADi of tests/golden/make_golden.py, builds the
map of path_generation/main.py:21-49 and times Problem.get_cost / get_nonlincon / Map.collides on arcs of
Solver.create_x_init (N = 80), one core -- the CPU path SURVEY.md 8(d)(1) asks to quote next to the GPU numbers.  The result is
written to profiles/ by hand (the reference cannot travel to the GPU box)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
import make_golden as mg  # noqa: E402


def main():
    n_paths = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    shim = tempfile.mkdtemp(prefix='uam_shim_')
    mg.build_shim(shim)
    sys.path.insert(0, mg.PG)
    sys.path.insert(0, shim)
    os.chdir(mg.PG)
    from region_map import RegionMap
    from problem import Problem
    from solver import Solver
    from ball import ball
    import utils as ut
    data = os.path.join(mg.REF, 'data', 'processed')
    m = RegionMap()
    m.add_obstacles(*[ball(c, r) for c, r in [([38.66652661075855, -9.203164091309498], 9), ([46.36137256675563, 3.9427562315386298], 2),
                                              ([19.846825121034392, 18.93411773399299], 2), ([26.037433469490207, 15.46710452712196], 2),
                                              ([46.87758543585609, -19.138710035318375], 2)]])
    m.new_region('Land', [0.9290, 0.6940, 0.1250])
    m.add_shapes_to_region('Land', *ut.get_var_from_file(os.path.join(data, 'land_area.txt'), 'vertices'))
    m.new_region('Population', 'Red')
    m.add_shapes_to_region('Population', *ut.get_var_from_file(os.path.join(data, 'populated_area.txt'), 'vertices'))
    m.new_region('HistCenter', 'Green')
    m.add_shape_to_region('HistCenter', ball([33.874752, -24.981154], 1))
    m.x_start, m.x_goal = [35.590685, -27.711422], [26.478673, 9.564082]
    N = 80
    pr = Problem(m, N, {'length_smooth': True, 'penalty_smooth': True, 'obstacle_smooth': True, 'maxratio_smooth': False})
    pr.params.update({'maxratio': 1.04, 'maxalpha': np.pi / 80, 'enlargement': 0})
    for name, w in zip(m.region_names(), [200, 15000, 27000]):
        pr.set_weight(name, w)
    sv = Solver(pr, {})
    Z = [np.concatenate([m.x_start, sv.create_x_init(d), m.x_goal]) for d in np.linspace(-0.9, 0.9, n_paths)]
    t0 = time.perf_counter()
    cost = [float(pr.get_cost(z)) for z in Z]
    t_cost = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = [np.asarray(pr.get_nonlincon(z), dtype=float) for z in Z]
    t_g = time.perf_counter() - t0
    t0 = time.perf_counter()
    col = [any(bool(m.collides(z[2 * j:2 * j + 2])) for j in range(N + 2)) for z in Z]
    t_col = time.perf_counter() - t0
    print(json.dumps({'what': 'reference Python (path_generation/problem.py, map.py) under the numpy stand-in for CasADi, 1 core, '
                              'main.py map (5 obstacles, 34 region shapes), N = 80',
                      'paths': n_paths, 'cores_in_container': os.cpu_count(),
                      'get_cost_s_per_path': t_cost / n_paths, 'get_nonlincon_s_per_path': t_g / n_paths,
                      'collides_s_per_path': t_col / n_paths,
                      'cost_and_collision_paths_per_s': n_paths / (t_cost + t_col),
                      'segment_evals_per_s': n_paths * (N + 1) / (t_cost + t_col),
                      'first_costs': cost[:3], 'collisions': int(sum(col))}))


if __name__ == '__main__':
    main()

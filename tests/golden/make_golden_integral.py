#!/usr/bin/env python
"""Reference values for the raster scorer's INTEGRAL mode (tests/golden/golden_integral.npz).

Not framework code.  Run only in the authoring container, where /root/reference exists:

    python tests/golden/make_golden_integral.py

The reference samples its penalty only at the waypoints (path_generation/problem.py:42-43); the line-integral mode
of the raster scorer is a build-defined extension:

    cost = (N+1) * length_of(z_) + (1/N) * ( sum_k mean_{s < S_k} P(z_k + s/S_k (z_{k+1} - z_k)) + P(z_{N+1}) )

This script pins it to the reference all the same: P is the REFERENCE'S OWN
`Problem.get_total_penalty_function()` (problem.py:49-82, run unmodified under the CasADi stand-in of
make_golden.py), evaluated at every sample position the scorer visits on a raster of R x R cells over the 64 km
window of the main.py map, and the length term is the reference's `length_of(z_, True)` (problem.py:130-146, the
call get_cost makes).  What is left between these values and the GPU's is the bilinear discretisation of the
rasterised field, which the parity test asserts to fall ~4x per doubling of R (tests/test_gpu_parity.py::
test_raster_integral_converges_to_reference).  The sample positions follow oracle/uam_oracle.py::score_paths_raster
(pixel coordinates, S_k = max(1, ceil(|dz_k|_cells * spc)), left-endpoint rule).
"""
import json
import multiprocessing as mp
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

WINDOW = (8.0, -42.0, 64.0)          # x0, y0, side (km): the window SURVEY.md 8(d) C1 names
RASTERS = (4096, 8192, 16384)
SPC = 1.0

_pen = None


def _setup():
    """the reference's map + problem from the committed fixture numbers, through the reference's own constructors"""
    shim = tempfile.mkdtemp(prefix='uam_shim_')
    mg.build_shim(shim)
    sys.path.insert(0, mg.PG)
    sys.path.insert(0, shim)
    os.chdir(mg.PG)
    from region_map import RegionMap
    from problem import Problem
    from polygon import polygon
    from ball import ball
    spec = json.load(open(os.path.join(HERE, 'fixture_main_map.json')))

    def mk(s):
        return polygon(*s['verts']) if s['kind'] == 'polygon' else ball(s['center'], s['r1'], s['r2'])
    m = RegionMap()
    m.add_obstacles(*[mk(s) for s in spec['obstacles']])
    for name, shapes in spec['regions']:
        m.new_region(name, 'red')
        m.add_shapes_to_region(name, *[mk(s) for s in shapes])
    m.x_start, m.x_goal = spec['x_start'], spec['x_goal']

    def problem(N):
        pr = Problem(m, N, dict(spec['options']))
        pr.params.update({'maxratio': spec['maxratio'], 'maxalpha': spec['maxalpha'], 'enlargement': spec['enlargement']})
        for name, w in zip(m.region_names(), spec['weights']):
            pr.set_weight(name, w)
        return pr
    return spec, problem


def _init_worker():
    global _pen
    _, problem = _setup()
    _pen = problem(5).get_total_penalty_function()       # P does not depend on N


def _eval(chunk):
    return [float(_pen(np.array(x))) for x in chunk]


def sample_positions(z_, R):
    """world coordinates of every sample of one path + per-segment counts (oracle/uam_oracle.py::score_paths_raster)"""
    x0, y0, side = WINDOW
    dx = side / R
    P = z_.reshape(-1, 2)
    U = (P[:, 0] - x0) / dx - 0.5
    V = (P[:, 1] - y0) / dx - 0.5
    dU, dV = U[1:] - U[:-1], V[1:] - V[:-1]
    S = np.maximum(1.0, np.ceil(np.sqrt(dU * dU + dV * dV) * SPC)).astype(np.int64)
    pts = []
    for k in range(len(S)):
        s = np.arange(S[k], dtype=np.float64)
        u, v = U[k] + s * (dU[k] / S[k]), V[k] + s * (dV[k] / S[k])
        pts.append(np.stack([x0 + (u + 0.5) * dx, y0 + (v + 0.5) * dx], axis=1))
    pts.append(P[-1:].copy())
    return np.concatenate(pts), S


def main():
    spec, problem = _setup()
    g = np.load(os.path.join(HERE, 'golden_ref.npz'))
    xs, xg = np.asarray(spec['x_start'], dtype=float), np.asarray(spec['x_goal'], dtype=float)
    paths = [('arc_N5', np.concatenate([xs, x, xg])) for x in g['arc_N5_x']] + \
            [('jit_N62', np.concatenate([xs, x, xg])) for x in g['jit_x'][:4]]
    out = {'window': np.array(WINDOW), 'rasters': np.array(RASTERS), 'spc': np.array(SPC)}
    out['paths_N5'] = np.stack([z for n, z in paths if n == 'arc_N5'])
    out['paths_N62'] = np.stack([z for n, z in paths if n == 'jit_N62'])
    pool = mp.Pool(os.cpu_count(), initializer=_init_worker)
    for R in RASTERS:
        costs = {'arc_N5': [], 'jit_N62': []}
        nsamp = {'arc_N5': [], 'jit_N62': []}
        for name, z_ in paths:
            N = len(z_) // 2 - 2
            pts, S = sample_positions(z_, R)
            chunks = np.array_split(pts, max(1, len(pts) // 256))
            vals = np.concatenate([np.asarray(v) for v in pool.map(_eval, chunks)])
            off = np.concatenate([[0], np.cumsum(S)])
            pen = sum(vals[off[k]:off[k + 1]].sum() / S[k] for k in range(len(S))) + vals[-1]
            L = float(problem(N).length_of(z_, True))           # the call get_cost makes (problem.py:39): quirk Q1 included
            costs[name].append((N + 1) * L + pen / N)
            nsamp[name].append(len(pts))
            print(R, name, len(pts), costs[name][-1], flush=True)
        for name in costs:
            out[f'cost_{name}_R{R}'] = np.array(costs[name])
            out[f'nsamples_{name}_R{R}'] = np.array(nsamp[name], dtype=np.int64)
    pool.close()
    np.savez_compressed(os.path.join(HERE, 'golden_integral.npz'), **out)
    print('wrote golden_integral.npz')


if __name__ == '__main__':
    main()

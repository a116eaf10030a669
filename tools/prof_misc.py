#!/usr/bin/env python
"""Small driver for ncu captures of the secondary kernels (analytic scorer with the shape grid, polygon front-end):
    ncu --set full --clock-control none -k regex:'uam_k_(score_analytic|ccl_merge|ccl_flatten|comp_stats|rect_extremes|rect_hull)' \
        -o gpurun_out/prof_misc python tools/prof_misc.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import uam_path_planning_b200 as uam
    rng = np.random.default_rng(20260101)
    # C2's analytic batch: 192 random boxes + 64 discs in one region, 32 obstacle discs, 10 000 arcs x 64 waypoints
    m = uam.RegionMap()
    m.new_region('Risk', 'r')
    for _ in range(192):
        c, a = rng.uniform(2, 62, 2), rng.uniform(0, np.pi)
        hw, hh = rng.uniform(0.25, 2.0, 2)
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        m.add_shape_to_region('Risk', uam.polygon(*(c + np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) @ R.T).tolist()))
    for _ in range(64):
        m.add_shape_to_region('Risk', uam.ball(rng.uniform(2, 62, 2).tolist(), float(rng.uniform(0.5, 3))))
    for _ in range(32):
        m.add_obstacle(uam.ball(rng.uniform(2, 62, 2).tolist(), float(rng.uniform(0.3, 1.5))))
    m.x_start, m.x_goal = [6.0, 7.0], [57.0, 55.0]
    prob = uam.Problem(m, 62, {})
    prob.params.update(maxratio=1.3, maxalpha=0.3, enlargement=0.0)
    prob.set_weight('Risk', 1000.0)
    sol = uam.Solver(prob, {})
    d = torch.rand(10000, device='cuda', dtype=torch.float64) * 1.8 - 0.9
    Z = sol.candidates_device(d)
    for _ in range(2):
        cost, col, _ = prob.score(Z)
    torch.cuda.synchronize()
    # polygon front-end on the 8192^2 land mask of the bench raster
    layers, occ, geo = bench.make_raster(torch, 'cuda:0')
    eng = uam.Engine()
    mask = (layers[0] > 0).to(torch.uint8)
    labels, n = eng.label_components(mask, 4)
    area, bbox = eng.component_stats(labels, n)
    ids = (torch.argsort(area, descending=True)[:16] + 1).to(torch.int32)
    rect = eng.component_rects(labels, n, bbox, ids, (0.0, 7.8125, 0.0, -7.8125))
    torch.cuda.synchronize()
    print('analytic best', float(cost.min()), 'components', n, 'largest', int(area.max()))


if __name__ == '__main__':
    main()

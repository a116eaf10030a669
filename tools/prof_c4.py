#!/usr/bin/env python
"""One pass of every map-rebuild kernel of config C4 (16384^2, 4096 rectangular footprints) for ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv python tools/prof_c4.py
    ncu --set full --clock-control none --import-source on -k regex:'uam_k_(rasterize_layers|layers_scan|occupancy_scan|cull|edt)' \
        -o gpurun_out/prof_c4 python tools/prof_c4.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import uam_path_planning_b200 as uam
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    modes = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1]
    rng = np.random.default_rng(20260104)
    KMn = 64.0 * n / 8192
    mm = uam.RegionMap()
    for r in ('Land', 'Population', 'Hist'):
        mm.new_region(r, 'r')
    for k in range(4096):
        c, a = rng.uniform(1, KMn - 1, 2), rng.uniform(0, np.pi)
        hw, hh = rng.uniform(0.45, 0.9, 2)
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        V = np.trunc((c + np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) @ R.T) * 1000) / 1000
        sh = uam.polygon(*V.tolist())
        mm.add_obstacle(sh)
        mm.add_shape_to_region(('Land', 'Population', 'Hist')[k % 3], sh)
    eng = uam.Engine()
    eng.set_shapes(mm.obstacles, mm._region_lists())
    geo = (0.0, KMn / n, 0.0, KMn / n)
    for mode in modes:
        eng.set_option('rasterizer', mode)
        occ = eng.rasterize_occupancy(n, n, geo)
        lay = eng.rasterize_layers(n, n, geo, 0.0)
        del lay
    d2, cl = eng.edt(occ, KMn / n)
    torch.cuda.synchronize()
    print('occupied', float(occ.float().mean()), 'max d2', int(d2.max()))


if __name__ == '__main__':
    main()

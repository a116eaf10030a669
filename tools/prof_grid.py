#!/usr/bin/env python
"""One batched cost-to-go sweep of config C5 (4096^2, random costs 1..999, 10 % blocked) for ncu:
    ncu --set full --clock-control none --import-source on -k regex:uam_k_grid_relax -s 200 -c 3 -o gpurun_out/prof_grid python tools/prof_grid.py [bands] [queries]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import uam_path_planning_b200 as uam
    bands = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    n = 4096
    g = torch.Generator(device='cuda').manual_seed(5)
    shape = (n, n) if bands == 1 else (bands, n, n)
    cost = torch.randint(1, 1000, shape, generator=g, device='cuda', dtype=torch.int32).to(torch.uint16)
    blk = (torch.rand(shape, generator=g, device='cuda') < 0.1).to(torch.uint8)
    rng = np.random.default_rng(11)
    src = [[int(v) for v in rng.integers(0, n, 2)] for _ in range(Q)]
    if bands > 1:
        src = [[int(rng.integers(0, bands))] + s for s in src]
    eng = uam.Engine()
    dist, _ = eng.grid_search(cost, src, blk, want_parent=False)
    torch.cuda.synchronize()
    print('activations', eng.get_stat('grid_activations'), 'sweeps', eng.get_stat('grid_sweeps'), 'rounds', eng.get_stat('grid_rounds'))


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Sweep of UAM_OPT_GRID_DELTA (width of the distance window relaxed per round) for the C5 start/goal queries.
    python tools/sweep_grid_delta.py > gpurun_out/sweep_grid_delta.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import uam_path_planning_b200 as uam
    dev = 'cuda'
    eng = uam.Engine()
    n5 = 4096
    out = []
    for bands, Q in ((1, 64), (8, 16)):
        g = torch.Generator(device=dev).manual_seed(5 + bands)
        shape = (n5, n5) if bands == 1 else (bands, n5, n5)
        cost = torch.randint(1, 1000, shape, device=dev, generator=g, dtype=torch.int32).to(torch.uint16)
        blk = (torch.rand(shape, device=dev, generator=g) < 0.1).to(torch.uint8)
        src = torch.randint(0, n5, (Q, 2), device=dev, generator=g, dtype=torch.int32)
        goal = torch.randint(0, n5, (Q, 2), device=dev, generator=g, dtype=torch.int32)
        if bands > 1:
            sb = torch.randint(0, bands, (Q, 1), device=dev, generator=g, dtype=torch.int32)
            src, goal = torch.cat([sb, src], dim=1), torch.cat([sb, goal], dim=1)
        blk[tuple(src[:, k].long() for k in range(src.shape[1]))] = 0
        ref = None
        for delta in (0, 16000, 32000, 64000, 128000, 256000, 512000, 1024000):
            eng.set_option('grid_delta', delta)
            best = 1e30
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                gd, path, plen = eng.grid_routes(cost, src, goal, blk, chunk=Q, max_len=8 * n5)
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
            if ref is None:
                ref = gd.clone()
            out.append({'bands': bands, 'queries': Q, 'delta': delta, 'seconds': best, 'queries_per_s': Q / best,
                        'activations': eng.get_stat('grid_activations'), 'rounds': eng.get_stat('grid_rounds'),
                        'same_goal_distances': bool(torch.equal(gd, ref))})
            print(json.dumps(out[-1]), flush=True)


if __name__ == '__main__':
    main()

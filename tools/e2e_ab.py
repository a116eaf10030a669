#!/usr/bin/env python
"""A/B of the host pipeline of uam_score_paths_raster_host on the bench workload (C3 shard):
    python tools/e2e_ab.py "chunks:taper" ...      e.g.  4:0 4:70 5:70 6:80
prints ms per call (wall clock around the call, pinned numpy in / out) for every setting."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import uam_path_planning_b200 as uam
    B = int(os.environ.get('AB_PATHS', '125000'))
    layers, occ, geo = bench.make_raster(torch, 'cuda:0')
    rm = uam.RasterMap.from_arrays(layers, geo, occ, device=0)
    Z = bench.make_paths(torch, 'cuda:0', B, 2000)
    Zh = torch.empty((B, 2 * bench.WP), dtype=torch.float64).pin_memory()
    Zh.copy_(Z)
    Zh = Zh.numpy()
    cost_h = torch.empty(B, dtype=torch.float32).pin_memory().numpy()
    col_h = torch.empty(B, dtype=torch.uint8).pin_memory().numpy()
    Zd = torch.empty_like(Z)
    Zt = torch.from_numpy(Zh)
    for _ in range(3):
        Zd.copy_(Zt, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        Zd.copy_(Zt, non_blocking=True)
    torch.cuda.synchronize()
    print(f'H2D of the {Zh.nbytes / 1e6:.0f} MB of waypoints alone: {(time.perf_counter() - t0) * 100:.3f} ms', flush=True)
    ref = None
    for cfg in sys.argv[1:]:
        ch, tp = (int(v) for v in cfg.split(':'))
        rm.engine.set_option('host_chunks', ch)
        rm.engine.set_option('host_taper', tp)
        for _ in range(3):
            rm.score_paths(Zh, bench.WEIGHTS, bench.SPC, True, None, out=(cost_h, col_h))
        ts = []
        for _ in range(10):
            t0 = time.perf_counter()
            rm.score_paths(Zh, bench.WEIGHTS, bench.SPC, True, None, out=(cost_h, col_h))
            ts.append((time.perf_counter() - t0) * 1e3)
        if ref is None:
            ref = cost_h.copy()
        assert np.array_equal(ref, cost_h), 'chunking changed the costs'
        print(f'chunks {ch} taper {tp:3d} %: median {np.median(ts):.3f} ms  min {np.min(ts):.3f} ms', flush=True)
    # the same call with an ordinary (pageable) numpy array, as a numpy user would make it
    rm.engine.set_option('host_chunks', 0)
    rm.engine.set_option('host_taper', 0)
    Zp = np.array(Zh, copy=True)
    cp, kp = np.empty(B, np.float32), np.empty(B, np.uint8)
    for _ in range(3):
        rm.score_paths(Zp, bench.WEIGHTS, bench.SPC, True, None, out=(cp, kp))
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        rm.score_paths(Zp, bench.WEIGHTS, bench.SPC, True, None, out=(cp, kp))
        ts.append((time.perf_counter() - t0) * 1e3)
    assert ref is None or np.array_equal(ref, cp)
    print(f'pageable numpy input, default chunks: median {np.median(ts):.3f} ms  min {np.min(ts):.3f} ms', flush=True)
    Zd2 = torch.empty_like(Z)
    t0 = time.perf_counter()
    for _ in range(5):
        Zd2.copy_(torch.from_numpy(Zp))
    torch.cuda.synchronize()
    print(f'(torch copy of the same pageable array to the device: {(time.perf_counter() - t0) * 200:.3f} ms)', flush=True)


if __name__ == '__main__':
    main()

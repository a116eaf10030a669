#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (one JSON line each; bench.py stays the headline):

  C2  10k candidate polylines (64 waypoints) on a 4096^2 risk+obstacle raster (L = 1), corridor and scatter
      distributions, waypoint and integral mode, plus the analytic scorer on the reference's main.py map
  C4  map rebuild at n^2 (default 16384): DEM mask, occupancy + 3 risk layers from 4096 rectangular footprints,
      exact EDT clearance -- Mcell/s per kernel and fraction of the HBM roofline (algorithmic bytes of SURVEY 8d)

    python bench_configs.py [--c4-size 16384] [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def ev_time(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run_c5(args, torch, uam, dev, rank=0, world=1, reduce_max=None):
    """C5 (BASELINE.json configs[4]): batched grid search on a 4096^2 grid x `bands` altitude bands of uint16 cost.
    Queries are independent: with `world` ranks the job's queries shard contiguously over the ranks
    (distributed.shard_range), the grid is replicated, there is no exchange at all; times are the max over ranks.
      full sweeps        c5_queries (1 band) / c5_queries_bands (8 bands) per GPU: whole distance + predecessor fields
      start/goal routes  c5_routes per GPU (128 x 8 GPUs = BASELINE's 1024 queries), in chunks of 64 / 16 through
                         Engine.grid_routes (bounded memory); the first chunk's goal distances must equal the full sweep's"""
    from uam_path_planning_b200 import distributed as udist
    eng = uam.Engine(torch.cuda.current_device())
    reduce_max = reduce_max or (lambda x: x)
    n5 = args.c5_size
    out = []
    for bands, Qf in ((1, args.c5_queries), (8, args.c5_queries_bands)):
        if Qf <= 0:
            continue
        Qr = max(Qf, getattr(args, 'c5_routes', Qf))
        g = torch.Generator(device=dev).manual_seed(5 + bands)          # same grid and query list on every rank
        shape = (n5, n5) if bands == 1 else (bands, n5, n5)
        cost = torch.randint(1, 1000, shape, device=dev, generator=g, dtype=torch.int32).to(torch.uint16)
        blk = (torch.rand(shape, device=dev, generator=g) < 0.1).to(torch.uint8)
        Qall = Qr * world
        src = torch.randint(0, n5, (Qall, 2), device=dev, generator=g, dtype=torch.int32)
        goal = torch.randint(0, n5, (Qall, 2), device=dev, generator=g, dtype=torch.int32)
        if bands > 1:
            sb = torch.randint(0, bands, (Qall, 1), device=dev, generator=g, dtype=torch.int32)
            src, goal = torch.cat([sb, src], dim=1), torch.cat([sb, goal], dim=1)
        idx = lambda t: tuple(t[:, k].long() for k in range(t.shape[1]))
        blk[idx(src)] = 0
        b0, b1 = udist.shard_range(Qall, rank, world)
        src, goal = src[b0:b1].contiguous(), goal[b0:b1].contiguous()
        Ql = b1 - b0
        # ---- full cost-to-go sweeps (distance + parent fields) of this rank's first Qf queries ---------------------
        dt = 1e30
        dist = parent = None
        for rep in range(1 + args.c5_reps):                             # first call = warm-up (scratch allocation)
            del dist, parent
            l0 = eng.launch_count()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dist, parent = eng.grid_search(cost, src[:Qf], blk)
            torch.cuda.synchronize()
            if rep:
                dt = min(dt, time.perf_counter() - t0)
        launches = eng.launch_count() - l0
        full_stats = {k: eng.get_stat('grid_' + k) for k in ('activations', 'sweeps', 'rounds', 'host_submissions')}
        reach = float((dist < 2 ** 62).float().mean().item())
        d_goal_full = dist[(torch.arange(Qf, device=dev),) + idx(goal[:Qf])].clone()
        cpu = None
        if not args.no_cpu and rank == 0 and (bands == 1 or getattr(args, 'c5_cpu_bands', True)):
            # CPU baseline: heap Dijkstra of oracle/uam_oracle_c.c, one query per thread, and a full-size
            # parity check of those queries
            from oracle import uam_oracle_c as occ
            cores = os.cpu_count() or 1
            nq = min(Qf, cores if bands == 1 else max(1, cores // 8))
            t0 = time.perf_counter()
            d_ref, _ = occ.grid_search(cost.cpu().numpy(), src[:nq].cpu().numpy(), blk.cpu().numpy(), want_parent=False, threads=cores)
            dtc = time.perf_counter() - t0
            cpu = {'queries': nq, 'threads': min(nq, cores), 'seconds': dtc, 'queries_per_s': nq / dtc,
                   'dist_equal_gpu': bool(np.array_equal(dist[:nq].cpu().numpy(), d_ref))}
            del d_ref
        del dist, parent
        # ---- start/goal routes of all of this rank's queries, 16 at a time ------------------------------------------
        dtg = 1e30
        chunk = 64 if bands == 1 else 16            # (chunk, bands, 4096, 4096) int64 + int32 fields: 12.9 GB / 25.8 GB
        for rep in range(1 + (args.c5_reps if bands == 1 else 0) + (1 if Qr <= 16 else 0)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gd, path, plen = eng.grid_routes(cost, src, goal, blk, chunk=chunk, max_len=8 * n5)
            torch.cuda.synchronize()
            if rep or bands > 1 and Qr > 16:
                dtg = min(dtg, time.perf_counter() - t0)
        goals_ok = bool(torch.equal(gd[:Qf], d_goal_full))
        mean_nodes, found = float(plen.float().mean().item()), int((plen > 0).sum().item())
        del gd, path, plen, cost, blk
        dt, dtg = reduce_max(dt), reduce_max(dtg)
        nodes = bands * n5 * n5
        edges = nodes * (8 + (2 if bands > 1 else 0))
        out.append({'config': f'C5: grid search on a {n5}^2 8-connected grid x {bands} altitude band(s), {world} B200: {Qall} start/goal '
                              f'queries ({Qr} per GPU), {Qf * world} full cost-to-go sweeps ({Qf} per GPU)', 'bands': bands, 'n_gpus': world,
                    'start_goal_queries': {'queries': Qall, 'seconds': dtg, 'queries_per_s': Qall / dtg,
                                           'goal_distances_equal_full_sweep': goals_ok, 'mean_path_nodes': mean_nodes,
                                           'paths_found_rank0': found, 'queries_per_launch': chunk},
                    'full_sweeps': {'queries': Qf * world, 'seconds': dt, 'queries_per_s': Qf * world / dt,
                                    'Mnode_per_s': Qf * world * nodes / dt / 1e6,
                                    'min_edge_relaxations_per_s': Qf * world * edges * reach / dt, 'kernels_run_per_call': launches,
                                    'reachable_fraction': reach, **{k + '_rank0': v for k, v in full_stats.items()}},
                    'cpu_baseline': cpu,
                    'note': 'exact distances + parents (bit-identical to Dijkstra); warp-per-tile Gauss-Seidel sweeps; the relaxation rounds are '
                            'looped on the device (CUDA graph WHILE node: host_submissions = what the host enqueued per call); queries '
                            'sharded over the ranks, grid replicated, no collective'})
    return out


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        return 6650.0


def run_c2(args, torch, uam, dev):
    peak = hbm_peak()
    rng = np.random.default_rng(20260101)
    # ---------------- C2 ------------------------------------------------------------------------------------
    n, KM, Wp, B = 4096, 64.0, 64, 10000
    m = uam.RegionMap()
    m.new_region('Risk', 'r')
    for _ in range(192):                                    # random boxes 0.5-4 km (minAreaRect-like) ...
        c, a = rng.uniform(2, 62, 2), rng.uniform(0, np.pi)
        hw, hh = rng.uniform(0.25, 2.0, 2)
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        m.add_shape_to_region('Risk', uam.polygon(*(c + np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) @ R.T).tolist()))
    for _ in range(64):                                     # ... and discs r in [0.5, 3] km
        m.add_shape_to_region('Risk', uam.ball(rng.uniform(2, 62, 2).tolist(), float(rng.uniform(0.5, 3))))
    for _ in range(32):
        m.add_obstacle(uam.ball(rng.uniform(2, 62, 2).tolist(), float(rng.uniform(0.3, 1.5))))
    m.x_start, m.x_goal = [6.0, 7.0], [57.0, 55.0]
    geo = (0.0, KM / n, 0.0, KM / n)
    t0 = time.perf_counter()
    rm = uam.RasterMap.from_map(m, n, n, geo, clearance=True)
    torch.cuda.synchronize()
    t_first = (time.perf_counter() - t0) * 1e3          # includes CUDA context creation, module load, first allocations
    del rm
    t0 = time.perf_counter()
    rm = uam.RasterMap.from_map(m, n, n, geo, clearance=True)
    torch.cuda.synchronize()
    t_build = (time.perf_counter() - t0) * 1e3          # shape upload + occupancy + layer + EDT + texel packing, wall clock
    prob = uam.Problem(m, Wp - 2, {'length_smooth': True, 'obstacle_smooth': True})
    prob.params.update(maxratio=1.04, maxalpha=np.pi / 80, enlargement=0)
    prob.set_weight('Risk', 5000.0)
    sol = uam.Solver(prob, {})
    cell = KM / n
    Zc = np.ascontiguousarray(sol.candidates(rng.uniform(-0.9, 0.9, B), jitter=0.25 * cell, rng=rng))   # corridor
    s, e = rng.uniform(0, KM, (B, 1, 2)), rng.uniform(0, KM, (B, 1, 2))
    Zs = (s + np.linspace(0, 1, Wp).reshape(1, Wp, 1) * (e - s) + rng.normal(0, 2 * cell, (B, Wp, 2))).reshape(B, 2 * Wp)
    out = {'config': 'C2: 10k polylines x 64 waypoints, 4096^2 risk+obstacle raster (L=1), 1 B200',
           'map_rebuild_ms_4096': t_build, 'first_call_ms_incl_cuda_init': t_first, 'unit': 'segment-evals/s'}
    for name, Zh in (('corridor', Zc), ('scatter', Zs)):
        Z = torch.from_numpy(np.ascontiguousarray(Zh)).to(dev)
        for mode, spc in (('waypoint', 0.0), ('integral', 1.0)):
            _, _, ns = rm.score_paths(Z, [5000.0], spc, want_nsamples=True)
            tot = int(ns.sum().item())
            ms = ev_time(torch, lambda: rm.score_paths(Z, [5000.0], spc), 20, 3)
            ab = (Wp - 1) * B * 16 + tot * (16 + 1) + B * 21
            out[f'{name}_{mode}'] = {'ms': ms, 'value': B * (Wp - 1) / (ms * 1e-3), 'samples': tot,
                                     'algorithmic_GBps': ab / (ms * 1e-3) / 1e9, 'frac_of_hbm_peak': ab / (ms * 1e-3) / 1e9 / peak}
        # end-to-end with host buffers
        t0 = time.perf_counter()
        for _ in range(10):
            rm.score_paths(Zh, [5000.0], 1.0)
        out[f'{name}_integral_e2e_ms'] = (time.perf_counter() - t0) * 100
    Z = torch.from_numpy(Zc).to(dev)
    ms = ev_time(torch, lambda: prob.score(Z), 10, 2)
    eng_a = prob.map.engine()
    eng_a.set_option('shape_grid', 0)
    ms_all = ev_time(torch, lambda: prob.score(Z), 10, 2)
    eng_a.set_option('shape_grid', 1)
    ms = ev_time(torch, lambda: prob.score(Z), 10, 2)
    ms_g = ev_time(torch, lambda: prob.score(Z, want_g=True), 10, 2)
    ms_grad = ev_time(torch, lambda: prob.get_cost_gradient(Z), 10, 2)
    out['analytic_corridor'] = {'ms': ms, 'value': B * (Wp - 1) / (ms * 1e-3), 'shapes': 256 + 32,
                                'ms_every_shape_at_every_point': ms_all, 'ms_with_constraint_vector': ms_g,
                                'ms_cost_and_gradient': ms_grad,
                                'shape_grid_cells': eng_a.get_stat('shape_grid_cells'),
                                'shape_grid_items': eng_a.get_stat('shape_grid_items'),
                                'note': 'fp64 analytic scorer (Problem.get_cost + collides) with per-cell candidate lists over the shapes'}
    del rm, Z
    return out


def run_c4(args, torch, uam, dev):
    peak = hbm_peak()
    rng = np.random.default_rng(20260104)
    # ---------------- C4 ------------------------------------------------------------------------------------
    n = args.c4_size
    sys.path.insert(0, ROOT)
    from bench import make_raster
    layers, occ0, _ = make_raster(torch, dev, n)
    dem = (layers[0] * 557.5).contiguous()
    dem[dem <= 0] = -9999.0                               # sea = nodata, like merge_test.tif
    del layers, occ0
    eng = uam.Engine()
    cells = n * n
    res = {'config': f'C4: map rebuild at {n}^2, 4096 rectangular footprints (integer-metre corners)', 'unit': 'Mcell/s',
           'hbm_peak_GBps': peak}
    ms = ev_time(torch, lambda: eng.dem_mask(dem, 0.0), args.reps)
    res['dem_mask'] = {'ms': ms, 'Mcell_s': cells / ms / 1e3, 'frac': cells * 5 / (ms * 1e-3) / 1e9 / peak}
    # polygon front-end: land mask -> 4-connected regions -> cell counts -> minimum-area rectangles of the
    # regions over min_area.  Wall-clock per call (each step reads a count back); algorithmic bytes of the labelling =
    # 1 B mask in + 4 B label out per cell
    mask = eng.dem_mask(dem, 0.0)
    cellm = 64000.0 * (n / 8192) / n                     # metres per cell (a 64 km map at 8192^2)
    geo_m = (0.0, cellm, 0.0, -cellm)

    def wall(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps, out
    ms_l, (labels, ncomp) = wall(lambda: eng.label_components(mask, 4))
    ms_s, (area, bbox) = wall(lambda: eng.component_stats(labels, ncomp))
    ids = (torch.nonzero(area.double() * cellm * cellm > 750000.0).reshape(-1) + 1).to(torch.int32)
    ms_r, rect = wall(lambda: eng.component_rects(labels, ncomp, bbox, ids, geo_m))
    res['polygon_front_end'] = {'components': ncomp, 'over_min_area': int(ids.numel()), 'label_ms': ms_l, 'stats_ms': ms_s,
                                'rects_ms': ms_r, 'label_Mcell_s': cells / ms_l / 1e3,
                                'label_frac_of_hbm_peak': cells * 5 / (ms_l * 1e-3) / 1e9 / peak}
    if not args.no_cpu:
        from oracle import uam_oracle as orc
        crop = mask[:4096, :4096].cpu().numpy()
        t0 = time.perf_counter()
        lab_c, n_c = orc.label_components(crop, 4)
        res['polygon_front_end']['cpu_scipy_label_Mcell_s'] = crop.size / (time.perf_counter() - t0) / 1e6
        lab_g, n_g = eng.label_components(mask[:4096, :4096].contiguous(), 4)
        res['polygon_front_end']['crop_labels_equal_scipy'] = bool(n_g == n_c and np.array_equal(lab_g.cpu().numpy(), lab_c))
    del dem, mask, labels, area, bbox, rect
    KMn = 64.0 * n / 8192
    mm = uam.RegionMap()
    for r in ('Land', 'Population', 'Hist'):
        mm.new_region(r, 'r')
    spec4 = {'obstacles': [], 'regions': [('Land', []), ('Population', []), ('Hist', [])]}
    for k in range(4096):
        c, a = rng.uniform(1, KMn - 1, 2), rng.uniform(0, np.pi)
        hw, hh = rng.uniform(0.45, 0.9, 2)                 # area > 0.78 km^2 like data_processor.py:9,32
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        V = np.trunc((c + np.array([[-hw, -hh], [hw, -hh], [hw, hh], [-hw, hh]]) @ R.T) * 1000) / 1000   # integer metres
        sh = uam.polygon(*V.tolist())
        mm.add_obstacle(sh)
        mm.add_shape_to_region(('Land', 'Population', 'Hist')[k % 3], sh)
        spec4['obstacles'].append({'kind': 'polygon', 'verts': V.tolist()})
        spec4['regions'][k % 3][1].append({'kind': 'polygon', 'verts': V.tolist()})
    eng.set_shapes(mm.obstacles, mm._region_lists())
    geo = (0.0, KMn / n, 0.0, KMn / n)
    occ = eng.rasterize_occupancy(n, n, geo)
    ms = ev_time(torch, lambda: eng.rasterize_occupancy(n, n, geo), args.reps, 1)
    res['occupancy'] = {'ms': ms, 'Mcell_s': cells / ms / 1e3, 'frac': cells * 1 / (ms * 1e-3) / 1e9 / peak,
                        'occupied_fraction': float(occ.float().mean().item())}
    ms = ev_time(torch, lambda: eng.rasterize_layers(n, n, geo, 0.0), args.reps, 1)
    res['risk_layers_x3'] = {'ms': ms, 'Mcell_s': cells / ms / 1e3, 'frac': cells * 12 / (ms * 1e-3) / 1e9 / peak}
    ms = ev_time(torch, lambda: eng.edt(occ, KMn / n), args.reps, 1)
    res['edt_clearance'] = {'ms': ms, 'Mcell_s': cells / ms / 1e3, 'frac': cells * (1 + 16 + 4) / (ms * 1e-3) / 1e9 / peak}
    tot = res['dem_mask']['ms'] + res['occupancy']['ms'] + res['risk_layers_x3']['ms'] + res['edt_clearance']['ms']
    res['total'] = {'ms': tot, 'Mcell_s': cells / tot / 1e3, 'frac': cells * 21 / (tot * 1e-3) / 1e9 / peak}
    # CPU baseline on a crop: scipy EDT (1 core)
    from scipy import ndimage
    crop = occ[:2048, :2048].cpu().numpy()
    t0 = time.perf_counter()
    ndimage.distance_transform_edt(crop == 0)
    res['cpu_edt_scipy_Mcell_s'] = crop.size / (time.perf_counter() - t0) / 1e6
    if not args.no_cpu:
        # CPU baseline of the rasterisation: the numpy oracle (vectorised half-plane tests over all shapes, one core) on a
        # 96 x 96-cell crop, which also checks the GPU cells of that crop
        from oracle import uam_oracle as orc
        om4 = orc.OMap(spec4)
        cr = 96
        t0 = time.perf_counter()
        occ_ref = orc.rasterize_occupancy(om4, cr, cr, *geo)
        res['cpu_occupancy_numpy_Mcell_s'] = cr * cr / (time.perf_counter() - t0) / 1e6
        res['crop_occupancy_equal_oracle'] = bool(np.array_equal(occ[:cr, :cr].cpu().numpy(), occ_ref))
        t0 = time.perf_counter()
        orc.rasterize_layers(om4, cr, cr, *geo, 0.0)
        res['cpu_layers_numpy_Mcell_s'] = cr * cr / (time.perf_counter() - t0) / 1e6
    del occ, mm, eng
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--c4-size', type=int, default=16384)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--skip-c4', action='store_true')
    ap.add_argument('--skip-c5', action='store_true')
    ap.add_argument('--c5-size', type=int, default=4096)
    ap.add_argument('--c5-queries', type=int, default=16)
    ap.add_argument('--c5-queries-bands', type=int, default=4)
    ap.add_argument('--c5-reps', type=int, default=2)
    ap.add_argument('--c5-routes', type=int, default=128, help='start/goal queries per GPU (128 x 8 GPUs = BASELINE config 5)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--only', default='', help="'c2' / 'c4' / 'c5': run only that config")
    args = ap.parse_args()
    import torch
    import uam_path_planning_b200 as uam
    assert torch.cuda.is_available()
    dev = 'cuda'
    if args.only in ('', 'c2'):
        print(json.dumps(run_c2(args, torch, uam, dev)))
    if args.only in ('', 'c4') and not args.skip_c4:
        print(json.dumps(run_c4(args, torch, uam, dev)))
    if args.only in ('', 'c5') and not args.skip_c5:
        for line in run_c5(args, torch, uam, dev):
            print(json.dumps(line))


if __name__ == '__main__':
    main()

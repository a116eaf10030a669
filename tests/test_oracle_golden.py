"""The oracle (oracle/uam_oracle.py) against golden vectors produced by the reference's own code.

Pins the CPU restatement before anything is compared with it (no GPU needed)."""
import math

import numpy as np
import pytest

from oracle import uam_oracle as orc
from conftest import full_paths

RTOL = 1e-13   # oracle vs reference float64 (observed ~1e-15; summation-order noise only)


@pytest.fixture(scope='module')
def omap(fixture_spec):
    return orc.OMap(fixture_spec)


def test_survey_appendix_b_values(golden, golden_meta):
    # SURVEY.md App. B.1 numbers reproduce from the reference run here
    assert golden['arc_N80_cost'][0] == pytest.approx(2497.8684368774984, rel=1e-14)
    assert golden['arc_N80_cost'][2] == pytest.approx(2565.281618963959, rel=1e-14)
    assert golden['arc_N62_cost'][3] == pytest.approx(2278.261899104732, rel=1e-14)
    assert golden_meta['survey_jitter_cost'] == pytest.approx(2660.2542351510133, rel=1e-14)
    assert golden['arc_N80_collide'].sum(axis=1).tolist() == [0, 0, 22, 34, 34]
    assert golden['arc_N80_g'].shape == (5, 650)


@pytest.mark.parametrize('N', [80, 62, 64, 5])
def test_arcs(omap, fixture_spec, golden, N):
    f = fixture_spec
    for i, d in enumerate(golden['arc_disp']):
        x = orc.create_x_init(f['x_start'], f['x_goal'], N, float(d))
        np.testing.assert_allclose(x, golden[f'arc_N{N}_x'][i], rtol=1e-15, atol=1e-15)
    X = golden[f'arc_N{N}_x']
    Z = full_paths(f, X)
    cost = orc.get_cost(omap, Z, N, f['weights'], f['enlargement'], f['options'])
    np.testing.assert_allclose(cost, golden[f'arc_N{N}_cost'], rtol=RTOL)
    g = orc.get_nonlincon(omap, Z, N, f['maxratio'], f['maxalpha'], f['options'])
    G = golden[f'arc_N{N}_g']
    np.testing.assert_allclose(g, G, rtol=1e-9, atol=1e-13)
    # zero pattern of the obstacle block is bit-exact (fp64, same op order)
    assert np.array_equal(g[:, 3 * N:] == 0, G[:, 3 * N:] == 0)
    np.testing.assert_allclose(orc.length_of(omap, X, N, False), golden[f'arc_N{N}_length'], rtol=RTOL)
    np.testing.assert_allclose(orc.length_of(omap, X, N, True), golden[f'arc_N{N}_length_smooth'], rtol=RTOL)
    col = omap.collides(Z.reshape(-1, 2)).reshape(Z.shape[0], N + 2)
    assert np.array_equal(col, golden[f'arc_N{N}_collide'])
    assert np.array_equal(orc.path_collides(omap, Z, N), golden[f'arc_N{N}_collide'].any(axis=1))


def test_jittered_paths(omap, fixture_spec, golden, golden_meta):
    f = fixture_spec
    N = 62
    Z = full_paths(f, golden['jit_x'])
    np.testing.assert_allclose(orc.get_cost(omap, Z, N, f['weights'], 0.0, f['options']), golden['jit_cost'], rtol=RTOL)
    g = orc.get_nonlincon(omap, Z, N, f['maxratio'], f['maxalpha'], f['options'])
    np.testing.assert_allclose(g, golden['jit_g'], rtol=1e-9, atol=1e-13)
    assert np.array_equal(g == 0, golden['jit_g'] == 0)
    col = omap.collides(Z.reshape(-1, 2)).reshape(-1, N + 2)
    assert np.array_equal(col, golden['jit_collide'])
    zs = full_paths(f, golden['survey_jitter_x'])
    c = orc.get_cost(omap, zs, N, f['weights'], 0.0, f['options'])
    assert c[0] == pytest.approx(golden_meta['survey_jitter_cost'], rel=RTOL)
    gs = orc.get_nonlincon(omap, zs, N, f['maxratio'], f['maxalpha'], f['options'])[0]
    np.testing.assert_allclose(gs, golden['survey_jitter_g'], rtol=1e-9, atol=1e-13)
    assert gs[2] == pytest.approx(0.037299322077608, rel=1e-9)


def test_variants(omap, fixture_spec, golden, golden_meta):
    f = fixture_spec
    N = 80
    z = full_paths(f, orc.create_x_init(f['x_start'], f['x_goal'], N, 0.0))
    v = golden_meta['variants_straight_N80']
    o = f['options']
    assert orc.get_cost(omap, z, N, f['weights'], 1.0, o)[0] == pytest.approx(v['enlargement1'], rel=RTOL)
    assert orc.get_cost(omap, z, N, f['weights'], -0.25, o)[0] == pytest.approx(v['enlargement_neg'], rel=RTOL)
    assert orc.get_cost(omap, z, N, f['weights'], 0.0, dict(o, length_smooth=False))[0] == pytest.approx(v['length_nonsmooth'], rel=RTOL)
    assert orc.get_cost(omap, z, N, [100, 7500, 13500], 0.0, o)[0] == pytest.approx(v['weights_alt'], rel=RTOL)
    assert math.isnan(v['penalty_nonsmooth'])                        # quirk Q4
    assert math.isnan(orc.get_cost(omap, z, N, f['weights'], 0.0, dict(o, penalty_smooth=False))[0])
    za = full_paths(f, golden['var_x'])
    g1 = orc.get_nonlincon(omap, za, N, f['maxratio'], f['maxalpha'], dict(o, obstacle_smooth=False))[0]
    np.testing.assert_allclose(g1, golden['var_g_obstacle_nonsmooth'], rtol=1e-9, atol=1e-13)
    g2 = orc.get_nonlincon(omap, za, N, f['maxratio'], f['maxalpha'], dict(o, maxratio_smooth=True))[0]
    np.testing.assert_allclose(g2, golden['var_g_maxratio_smooth'], rtol=1e-9, atol=1e-13)


def test_point_queries(omap, fixture_spec, golden):
    f = fixture_spec
    Q = golden['pt_x']
    np.testing.assert_allclose(orc.total_penalty(omap, Q, f['weights'], 0.0, True), golden['pt_total_penalty'], rtol=RTOL, atol=0)
    for l, ((name, shapes), w) in enumerate(zip(omap.regions, f['weights'])):
        np.testing.assert_allclose(orc.region_penalty(shapes, Q, w, True, 0.0), golden['pt_region_penalty'][:, l], rtol=RTOL)
    np.testing.assert_allclose(orc.obstacle_penalty(omap, Q, 0.0, True), golden['pt_obstacle_penalty'], rtol=RTOL)
    assert np.array_equal(omap.collides(Q), golden['pt_collides'])
    assert np.array_equal(golden['pt_collides'], golden['pt_getitem'])
    # SURVEY B.2 spot values
    assert golden['pt_total_penalty'][0] == pytest.approx(29616.583973980643, rel=1e-14)
    assert golden['pt_obstacle_penalty'][6] == pytest.approx(0.9555804225026664, rel=1e-14)


def test_inequalities_bit_exact(omap, golden):
    """h_i(x) of every shape at 64 points: identical bits -> same edge order, sign and op order."""
    shapes = list(omap.obstacles) + [s for _, ss in omap.regions for s in ss]
    off = golden['h_offsets']
    P = golden['h_points']
    for k, s in enumerate(shapes):
        H = s.h(P)
        assert H.shape[0] == off[k + 1] - off[k]
        assert np.array_equal(H, golden['h_values'][off[k]:off[k + 1]]), k
        np.testing.assert_array_equal(s.center, golden['shape_centers'][k])
        assert s.area == pytest.approx(golden['shape_areas'][k], rel=1e-14)
        assert s.psi(s.center.reshape(1, 2))[0] == pytest.approx(golden['shape_psi_center'][k], rel=RTOL)
    sp = orc.make_polygon(golden['shuffled_poly_verts'])
    assert np.array_equal(sp.h(P), golden['shuffled_poly_h'])
    np.testing.assert_array_equal(sp.center, golden['shuffled_poly_center'])


def test_constructors(golden_meta):
    c = golden_meta['constructors']
    sq = orc.make_square([1, 1], 0.5)
    bl = orc.make_ball([1, 1], 2, 1)
    assert bool(sq.contains([1.2, 0.9])[0]) == c['square_contains_1.2_0.9']
    assert bool(sq.contains([1.6, 1.0])[0]) == c['square_contains_1.6_1']
    assert sq.psi([1.2, 0.9])[0] == pytest.approx(c['square_psi_1.2_0.9'], rel=1e-15)
    assert bool(bl.contains([2.9, 1.0])[0]) == c['ball_contains_2.9_1']
    assert bool(bl.contains([1.0, 2.1])[0]) == c['ball_contains_1_2.1']
    us = orc.make_polygon([[0., 0.], [1., 0.], [1., 1.], [0., 1.]])
    assert bool(us.contains([1 + 1e-15, .5])[0]) == c['unit_square_edge_1e-15'] is True
    assert bool(us.contains([1 + 1e-13, .5])[0]) == c['unit_square_edge_1e-13'] is False
    for name, pts in {'two_vertices': [[0., 0.], [1., 1.]], 'aligned': [[0., 0.], [1., 0.], [2., 0.], [1., 1.]],
                      'nonconvex': [[0., 0.], [2., 0.], [0.5, 0.5], [0., 2.]]}.items():
        with pytest.raises(ValueError) as ei:
            orc.make_polygon(pts)
        assert str(ei.value) == c['errors'][name][1]
    with pytest.raises(ValueError) as ei:
        orc.create_x_init([0, 0], [1, 1], 10, 1.5)
    assert str(ei.value) == c['errors']['x_init_1.5'][1]


def test_testscript_config1():
    """tests/test_path_generation.py inline problem at its own initial guess."""
    z0, zN = np.array([35.590685, -27.711422]), np.array([26.478673, 9.564082])
    c = np.array([31.034679, -9.07367])
    z = np.linspace(z0, zN, 6)[1:-1]
    dist, pen, tot = orc.testscript_cost(z, z0, zN, c)
    assert dist == pytest.approx(294.498392228432, rel=1e-13)
    assert pen == 0 and tot == pytest.approx(294.498392228432, rel=1e-13)
    assert np.all(orc.testscript_constraints(z, z0, zN) < 1e-12) and len(orc.testscript_constraints(z, z0, zN)) == 9
    z2 = z.copy()
    z2[2] = c + 0.5
    assert orc.testscript_cost(z2, z0, zN, c)[1] == pytest.approx(2.25, rel=1e-12)


def test_raster_reduces_to_analytic(omap, fixture_spec, golden):
    """Raster waypoint mode ~ analytic reference cost (discretisation only): the tie of the
    build-defined raster formulation back to the reference (7e-5 @ 62.5 m cells)."""
    f = fixture_spec
    H = W = 512
    x0, y0, dx = 8.0, -42.0, 64.0 / W
    layers = orc.rasterize_layers(omap, H, W, x0, dx, y0, dx, 0.0)
    occ = orc.rasterize_occupancy(omap, H, W, x0, dx, y0, dx)
    N = 62
    Z = full_paths(f, golden['jit_x'])
    cost, col, ns = orc.score_paths_raster(layers, occ, (x0, dx, y0, dx), Z, f['weights'], 0.0, True, f['x_start'])
    ref = golden['jit_cost']
    assert np.max(np.abs(cost - ref) / ref) < 1e-2          # 125 m cells (observed 3.5e-3)
    assert np.all(ns == N + 2)
    # cell-step integral mode is a refinement of the same functional: stays close, uses more samples
    cost2, col2, ns2 = orc.score_paths_raster(layers, occ, (x0, dx, y0, dx), Z[:4], f['weights'], 1.0, True, f['x_start'])
    assert np.all(ns2 > ns[:4])
    assert np.max(np.abs(cost2 - ref[:4]) / ref[:4]) < 0.2
    assert np.all(col2 >= col[:4])


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - b) / np.abs(b)))


def test_raster_converges_to_reference_waypoint_mode(omap, fixture_spec, golden):
    """Raster waypoint mode against the REFERENCE'S analytic get_cost (golden jit_cost, problem.py:38-44) at 2048^2,
    4096^2 and 8192^2 over the 64 km window: the error is the bilinear discretisation of the rasterised penalty field and
    nothing else -- second order in the cell size (x ~3.4-4 per doubling: 1.7e-4, 5.1e-5, 1.5e-5; 3.6e-6 at 16384^2,
    asserted on the GPU where the full rasterisation is cheap).  float32 vs float64 texels make no difference at this
    level (checked when the rate was measured).  Only the texels the paths touch are rasterised here."""
    f = fixture_spec
    Z = full_paths(f, golden['jit_x'])
    ref = golden['jit_cost']
    err = {}
    for R in (2048, 4096, 8192):
        geo = (8.0, 64.0 / R, -42.0, 64.0 / R)
        layers, occ = orc.rasterize_along_paths(omap, Z, R, R, *geo, 0.0)
        cost, col, ns = orc.score_paths_raster(layers, occ, geo, Z, f['weights'], 0.0, True, f['x_start'])
        err[R] = _rel(cost, ref)
        del layers, occ
    assert err[2048] < 2.5e-4 and err[4096] < 7e-5 and err[8192] < 2e-5, err
    assert err[2048] / err[4096] > 2.8 and err[4096] / err[8192] > 2.8, err


def test_raster_integral_mode_pinned_to_reference(omap, fixture_spec):
    """Integral mode against golden_integral.npz = the reference's own get_total_penalty_function (problem.py:49-82) at
    every sample position of the line integral + the reference's length_of (tests/golden/make_golden_integral.py).  Same
    sample counts; cost error = bilinear discretisation, x 4 per doubling: 3.9e-5 at 4096^2, 9.7e-6 at 8192^2 (7.8 m
    cells -- the cell size of the benchmarked configuration), i.e. within north_star's 1e-5 at the benchmarked size."""
    import os
    gi = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'golden_integral.npz'))
    f = fixture_spec
    x0, y0, side = gi['window']
    err = {}
    for R in (4096, 8192):
        geo = (x0, side / R, y0, side / R)
        e = 0.0
        for name, key in (('N5', 'arc_N5'), ('N62', 'jit_N62')):
            Z = gi['paths_' + name]
            layers, occ = orc.rasterize_along_paths(omap, Z, R, R, *geo, float(gi['spc']))
            cost, col, ns = orc.score_paths_raster(layers, occ, geo, Z, f['weights'], float(gi['spc']), True, f['x_start'])
            assert np.array_equal(ns, gi[f'nsamples_{key}_R{R}'])
            e = max(e, _rel(cost, gi[f'cost_{key}_R{R}']))
            del layers, occ
        err[R] = e
    assert err[4096] < 5e-5 and err[8192] <= 1e-5, err
    assert err[4096] / err[8192] > 3.0, err


def test_oracle_gradient_matches_finite_differences(omap, fixture_spec, golden):
    """The analytic gradient restatement against central differences of the golden-pinned get_cost."""
    f = fixture_spec
    N = 62
    Z = full_paths(f, golden['jit_x'][:3])
    for opts, e in [(f['options'], 0.0), (dict(f['options'], length_smooth=False), 0.3)]:
        G = orc.get_cost_gradient(omap, Z, N, f['weights'], e, opts)
        rng = np.random.default_rng(0)
        for col in rng.choice(2 * (N + 2), 24, replace=False):
            if col < 2 and not opts['length_smooth']:
                continue          # |z_0 - map.x_start| has a kink at 0: no derivative there (the zero subgradient is returned)
            h = 1e-6
            Zp, Zm = Z.copy(), Z.copy()
            Zp[:, col] += h
            Zm[:, col] -= h
            fd = (orc.get_cost(omap, Zp, N, f['weights'], e, opts) - orc.get_cost(omap, Zm, N, f['weights'], e, opts)) / (2 * h)
            np.testing.assert_allclose(G[:, col], fd, rtol=2e-5, atol=2e-4)
    assert np.max(np.abs(G)) > 1.0 and np.isfinite(G).all()       # the (map.x_start, z_0) kink contributes 0, not 0/0


# ------------------------------------------------------------------------------------------------------------
# the C / OpenMP form of the oracle (oracle/uam_oracle_c.c) against the numpy oracle it restates
# ------------------------------------------------------------------------------------------------------------
def _c_oracle():
    from oracle import uam_oracle_c as occ
    occ.build()
    return occ


@pytest.mark.parametrize('L', [1, 3])
def test_c_oracle_raster_matches_numpy_oracle(L):
    occ = _c_oracle()
    rng = np.random.default_rng(40 + L)
    H, W, geo = 90, 140, (3.0, 0.25, 40.0, -0.2)
    lay = rng.uniform(0, 2, (L, H, W)).astype(np.float32)
    oc = (rng.uniform(size=(H, W)) < 0.04).astype(np.uint8)
    B, Wp = 60, 9
    org, ext = np.array([3.0, 40.0]), np.array([0.25 * W, -0.2 * H])
    s = org + rng.uniform(-0.15, 1.15, (B, 1, 2)) * ext          # some paths start / end outside the raster
    g = org + rng.uniform(-0.15, 1.15, (B, 1, 2)) * ext
    t = np.linspace(0, 1, Wp).reshape(1, Wp, 1)
    Z = (s + t * (g - s) + rng.normal(0, 0.4, (B, Wp, 2))).reshape(B, 2 * Wp)
    Z[5, 4:] = np.tile(Z[5, 2:4], Wp - 2)                        # zero-length segments
    w = [200.0, 15000.0, 27000.0][:L]
    for spc in (0.0, 1.0, 0.37, 2.5):
        for ls in (True, False):
            for xs in (None, [4.0, 39.0]):
                a = orc.score_paths_raster(lay, oc, geo, Z, w, spc, ls, xs)
                for threads in (1, 3):
                    b = occ.score_paths_raster(lay, oc, geo, Z, w, spc, ls, xs, threads=threads)
                    np.testing.assert_allclose(b[0], a[0], rtol=1e-13, atol=0)
                    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_c_oracle_grid_search_matches_numpy_oracle():
    occ = _c_oracle()
    rng = np.random.default_rng(8)
    cost = rng.integers(1, 3000, (3, 17, 23)).astype(np.uint16)
    cost[rng.uniform(size=cost.shape) < 0.05] = 65535
    blk = (rng.uniform(size=cost.shape) < 0.15).astype(np.uint8)
    srcs = [[0, 3, 4], [2, 10, 10], [1, 16, 22], np.argwhere(blk == 1)[0].tolist()]
    d, p = occ.grid_search(cost, srcs, blk)
    for q, sv in enumerate(srcs):
        dr, pr = orc.grid_search(cost, tuple(sv), blk)
        assert np.array_equal(d[q], dr) and np.array_equal(p[q].astype(np.int64), pr)
    assert (d[3] == 2 ** 62).all() and (p[3] == -1).all()
    d2, p2 = occ.grid_search(cost[1], [[5, 6]], blk[1])
    dr, pr = orc.grid_search(cost[1], (5, 6), blk[1])
    assert np.array_equal(d2[0], dr) and np.array_equal(p2[0].astype(np.int64), pr)
    d3, _ = occ.grid_search(cost[1], [[5, 6]], None, want_parent=False)
    assert np.array_equal(d3[0], orc.grid_search(cost[1], (5, 6), None)[0])


def test_oracle_components_and_min_area_rect():
    """Oracle of the polygon front-end: labels number the regions in raster-scan order; the exact minimum-area rectangle
    has the area cv2.minAreaRect finds (the reference's call, map_generation/data_processor.py:67-71) and is never larger
    than any rectangle aligned with another direction."""
    mask = np.array([[1, 1, 0, 0, 1],
                     [0, 1, 0, 1, 1],
                     [0, 0, 1, 0, 0],
                     [1, 0, 1, 1, 0]], dtype=np.uint8)
    lab4, n4 = orc.label_components(mask, 4)
    assert n4 == 4 and lab4[0, 0] == 1 and lab4[0, 4] == 2 and lab4[2, 2] == 3 and lab4[3, 0] == 4 and lab4[1, 3] == 2
    lab8, n8 = orc.label_components(mask, 8)
    assert n8 == 2 and lab8[2, 2] == 1 and lab8[1, 3] == 1 and lab8[3, 0] == 2
    area, bbox = orc.component_stats(lab4, n4)
    assert area.tolist() == [3, 3, 3, 1] and bbox[1].tolist() == [0, 1, 3, 4]
    # unit cell, axis-aligned box, a tilted lattice parallelogram
    c, a, nh, e = orc.min_area_rect_exact([(0, 0), (1, 0), (0, 1), (1, 1)])
    assert a == 1 and nh == 4
    c, a, nh, e = orc.min_area_rect_exact([(x, y) for x in range(5) for y in range(3)])
    assert a == 8
    c, a, nh, e = orc.min_area_rect_exact([(0, 0), (3, 1), (4, 4), (1, 3)])
    assert float(a) < 16 and nh == 4
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(1)
    for _ in range(40):
        pts = rng.integers(0, 60, (rng.integers(3, 40), 2))
        if len(orc._hull_int(pts)) < 3:
            continue
        c, a, nh, e = orc.min_area_rect_exact(pts)
        (_, (w, h), _) = cv2.minAreaRect(pts.astype(np.float32))
        assert abs(float(a) - w * h) <= 2e-5 * w * h + 1e-3
        # corners form a rectangle of that area
        d1, d2 = c[1] - c[0], c[3] - c[0]
        assert abs(d1 @ d2) < 1e-9 * (1 + abs(d1).max() * abs(d2).max())
        assert abs(np.linalg.norm(d1) * np.linalg.norm(d2) - float(a)) < 1e-9 * (1 + float(a))


def test_oracle_large_polygon_split():
    """DataProcessor._divide_and_approximate_polygon (map_generation/data_processor.py:34-53) on the raster: a solid
    rectangle splits into divisions^2 equal boxes (the box edges fall inside cells: 7 x 11 cells / 5); a ring's hole cuts
    boxes into several pieces; every piece's rectangle has the area cv2.minAreaRect gives for the clipped cells' corners
    and the pieces' areas add up to at least the component's."""
    geo = (100.0, 2.0, 50.0, -2.0)
    lab = np.zeros((20, 30), dtype=np.int32)
    lab[3:10, 4:15] = 7                                   # 7 rows x 11 columns
    r, b = orc.split_component_rects(lab, 7, (3, 9, 4, 14), geo, 5)
    assert len(r) == 25 and sorted(b.tolist()) == list(range(25))
    bw, bh = 11 * 2.0 / 5, 7 * 2.0 / 5
    for rect, box in zip(r, b):
        j, k = divmod(int(box), 5)                        # x index outer; k counts from miny: y decreases with the row here
        xs, ys = rect[:, 0], rect[:, 1]
        assert abs(xs.min() - (100.0 + 4 * 2.0 + j * bw)) < 1e-9 and abs(xs.max() - xs.min() - bw) < 1e-9
        assert abs(ys.min() - (50.0 - 10 * 2.0 + k * bh)) < 1e-9 and abs(ys.max() - ys.min() - bh) < 1e-9
    # a comb: a bar with teeth two cells wide every three cells -- every box below the bar holds two separate pieces
    comb = np.zeros((40, 40), dtype=np.int32)
    comb[5:9, 5:35] = 1
    for c in range(5, 35, 3):
        comb[9:35, c:c + 2] = 1
    r, b = orc.split_component_rects(comb, 1, (5, 34, 5, 34), (0.0, 1.0, 0.0, 1.0), 5)
    counts = np.bincount(b, minlength=25).reshape(5, 5)          # [x box][y box]
    assert (counts[:, 0] == 1).all() and (counts[:, 1:] == 2).all() and len(r) == 45
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(9)
    f = rng.random((60, 70))
    for _ in range(6):
        f = (f + np.roll(f, 1, 0) + np.roll(f, -1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 1)) / 5.0
    labels, n = orc.label_components(f > np.quantile(f, 0.45), 4)
    area, bbox = orc.component_stats(labels, n)
    cid = int(np.argmax(area)) + 1
    r, b = orc.split_component_rects(labels, cid, bbox[cid - 1], (0.0, 1.0, 0.0, 1.0), 5)
    assert uam_rect_area(r).sum() >= area[cid - 1] - 1e-9
    r0, r1, c0, c1 = bbox[cid - 1]
    nr, nc = r1 - r0 + 1, c1 - c0 + 1
    fine = np.kron((labels[r0:r1 + 1, c0:c1 + 1] == cid).astype(np.uint8), np.ones((5, 5), dtype=np.uint8))
    k = 0
    for j in range(5):
        for kk in range(5):
            sub, ns = orc.label_components(fine[kk * nr:(kk + 1) * nr, j * nc:(j + 1) * nc], 4)
            for p in range(1, ns + 1):
                ii, jj = np.nonzero(sub == p)
                pts = np.concatenate([np.stack([jj + a, ii + c], 1) for a in (0, 1) for c in (0, 1)]).astype(np.float32)
                (_, (w, h), _) = cv2.minAreaRect(pts)
                assert abs(uam_rect_area(r[k:k + 1])[0] * 25 - w * h) <= 2e-5 * w * h + 1e-3
                k += 1
    assert k == len(r)


def uam_rect_area(rect):
    x, y = rect[..., 0], rect[..., 1]
    return 0.5 * np.abs(np.sum(x * np.roll(y, -1, axis=-1) - np.roll(x, -1, axis=-1) * y, axis=-1))

"""Float64 numpy restatement of the reference hot path -- TEST INFRASTRUCTURE ONLY.

This module is the *oracle*: a CPU restatement of nomaporon/uam_path_planning's
per-path cost / constraint / collision arithmetic, written from the reference's
behaviour (file:line citations below are relative to
``/root/reference/geo_simulation_project/path_generation``).  It is pinned by
``tests/golden/*.json|npz`` which were generated in the authoring container by
running the reference's OWN Python (``problem.py``, ``quadratic_obstacle.py``,
``polygon.py`` ...) under a numpy stand-in for CasADi
(``tests/golden/make_golden.py``).  The reference ships no golden vectors or
assertions of its own for this path, so that is the pin.

Sections whose semantics are BUILD-DEFINED EXTENSIONS (no counterpart in the
reference, "parity unpinned" by the reference): the raster formulation
(``rasterize_*``, ``score_paths_raster``), ``edt`` and ``grid_search``.  The
raster scorer is tied back to the reference by construction: with one sample per
segment it evaluates the reference's cost formula with the analytic penalty
replaced by a bilinear lookup of that same penalty rasterised at cell centres.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this file.  The product never does.

Shape specs (plain dicts, independent of the product's classes):
    {'kind': 'polygon', 'verts': [[x, y], ...]}
    {'kind': 'ball',    'center': [cx, cy], 'r1': r1, 'r2': r2}
    {'kind': 'square',  'center': [cx, cy], 'r1': r1, 'r2': r2}
Map spec:
    {'obstacles': [spec, ...], 'regions': [(name, [spec, ...]), ...],
     'x_start': [x, y], 'x_goal': [x, y]}
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

F64 = np.float64


# --------------------------------------------------------------------------- #
# shapes  (polygon.py:7-143, ball.py:7-52, square.py:6-65)
# --------------------------------------------------------------------------- #
class OShape:
    """Ordered list of inequalities h_i(x) <= 0 plus the reference's `center`."""

    def __init__(self, kind: str, edges: List[tuple], center, area: float):
        self.kind = kind
        self.edges = edges          # list of ('line', Ax, Ay, Bx, By, sgn) | ('ellipse', cx, cy, r1, r2) | ('box', axis, sign, c, r)
        self.center = None if center is None else np.asarray(center, dtype=F64).reshape(2)
        self.area = area

    # h_i at M points -> (n_edges, M); operation order as in the reference so that
    # `contains` is bit-exact (polygon.py:69-71,98; ball.py:33-37; square.py:29-51).
    def h(self, X: np.ndarray) -> np.ndarray:
        X = np.asarray(X, dtype=F64).reshape(-1, 2)
        x0, x1 = X[:, 0], X[:, 1]
        out = np.empty((len(self.edges), X.shape[0]), dtype=F64)
        for i, e in enumerate(self.edges):
            if e[0] == 'line':
                _, Ax, Ay, Bx, By, sgn = e
                line = (By - Ay) * (x0 - Ax) - (Bx - Ax) * (x1 - Ay)
                out[i] = -sgn * line
            elif e[0] == 'ellipse':
                _, cx, cy, r1, r2 = e
                a = (x0 - cx) / r1
                b = (x1 - cy) / r2
                out[i] = (a * a + b * b) - 1.0
            elif e[0] == 'box':
                _, axis, sign, c, r = e
                xd = x0 if axis == 0 else x1
                if sign > 0:          # x_d - c_d - r_d          (square.py:29,41)
                    out[i] = xd - c - r
                else:                 # -x_d + c_d - r_d         (square.py:35,47)
                    out[i] = -xd + c - r
            else:  # pragma: no cover
                raise ValueError(e[0])
        return out

    def contains(self, X) -> np.ndarray:
        """all_i h_i(x) <= 1e-14   (quadratic_obstacle.py:89-94)."""
        return np.all(self.h(X) <= 1e-14, axis=0)

    def psi(self, X, smooth: bool = True, e: float = 0.0) -> np.ndarray:
        """prod_i min(h_i - e, 0)^2 (smooth) | prod_i min(e - h_i, 0)   (quadratic_obstacle.py:27-39)."""
        H = self.h(X)
        res = np.ones(H.shape[1], dtype=F64)
        for i in range(H.shape[0]):
            if smooth:
                res = res * np.minimum(H[i] - e, 0.0) ** 2
            else:
                res = res * np.minimum(e - H[i], 0.0)
        return res


def _line_F(A, B, P):
    return (B[1] - A[1]) * (P[0] - A[0]) - (B[0] - A[0]) * (P[1] - A[1])


def make_polygon(verts: Sequence[Sequence[float]]) -> OShape:
    """Gift-wrap ordering from vertex 0 + convexity check   (polygon.py:20-21,55-136)."""
    if len(verts) < 3:
        raise ValueError(f'Only {len(verts)} vertices given. At least 3 required')
    pts = [np.asarray(p, dtype=F64).reshape(2) for p in verts]
    n = len(pts)
    center = pts[0].copy()                       # polygon.py:32-37 (sum in input order)
    for b in range(1, n):
        center = center + pts[b]

    def are_consecutive(a, b):
        sgn = 0.0
        for j in range(n):
            if j == a or j == b:
                continue
            s1 = np.sign(_line_F(pts[a], pts[b], pts[j]))
            if s1 == 0:
                raise ValueError('Input contains three aligned points')
            if sgn == 0:
                sgn = s1
                continue
            if s1 != sgn:
                return False, None
        if sgn == 0:
            raise ValueError('The polygon is nonconvex')
        return True, ('line', pts[a][0], pts[a][1], pts[b][0], pts[b][1], float(sgn))

    edges = []
    remaining = list(range(1, n))
    a = 0
    area = 0.0
    while remaining:
        found = False
        for i, b in enumerate(remaining):
            ok, f = are_consecutive(a, b)
            if ok:
                area += pts[a][0] * pts[b][1] - pts[a][1] * pts[b][0]
                remaining.pop(i)
                a = b
                edges.append(f)
                found = True
                break
        if not found:
            raise ValueError('The polygon is nonconvex')
    ok, f = are_consecutive(a, 0)
    if not ok:
        raise ValueError("Couldn't close polygon")
    area += pts[a][0] * pts[0][1] - pts[a][1] * pts[0][0]
    edges.append(f)
    return OShape('polygon', edges, center / n, abs(area) / 2)


def make_ball(center, r1=None, r2=None) -> OShape:
    """ball.py:19-24,33-37,49-50."""
    if r1 is None and r2 is None:
        r1 = center
        r2 = r1
        center = [0.0, 0.0]
    elif r2 is None:
        r2 = r1
    c = np.asarray(center, dtype=F64)
    assert c.shape == (2,)
    return OShape('ball', [('ellipse', c[0], c[1], float(r1), float(r2))], c, math.pi * r1 * r2)


def make_square(center, r1, r2=None) -> OShape:
    """square.py:18-51: sides right, left, top, bottom."""
    c = np.asarray(center, dtype=F64).reshape(2)
    if r2 is None:
        r2 = r1
    edges = [('box', 0, +1, c[0], float(r1)), ('box', 0, -1, c[0], float(r1)),
             ('box', 1, +1, c[1], float(r2)), ('box', 1, -1, c[1], float(r2))]
    return OShape('square', edges, c, 4 * r1 * r2)


def make_shape(spec: Dict) -> OShape:
    k = spec['kind']
    if k == 'polygon':
        return make_polygon(spec['verts'])
    if k == 'ball':
        return make_ball(spec['center'], spec.get('r1'), spec.get('r2'))
    if k == 'square':
        return make_square(spec['center'], spec['r1'], spec.get('r2'))
    raise ValueError(k)


class OMap:
    """Hard obstacles + ordered named regions   (map.py:7-17, region_map.py:8-61)."""

    def __init__(self, spec: Dict):
        self.obstacles = [make_shape(s) for s in spec.get('obstacles', [])]
        self.regions: List[Tuple[str, List[OShape]]] = [
            (name, [make_shape(s) for s in shapes]) for name, shapes in spec.get('regions', [])]
        self.x_start = np.asarray(spec.get('x_start', [0.0, 0.0]), dtype=F64)
        self.x_goal = np.asarray(spec.get('x_goal', [0.0, 0.0]), dtype=F64)

    def collides(self, X) -> np.ndarray:
        """any_o contains_o(x)   (map.py:41-43)."""
        X = np.asarray(X, dtype=F64).reshape(-1, 2)
        out = np.zeros(X.shape[0], dtype=bool)
        for o in self.obstacles:
            out |= o.contains(X)
        return out


DEFAULT_OPTIONS = {'length_smooth': False, 'penalty_smooth': True,
                   'obstacle_smooth': False, 'maxratio_smooth': False}   # problem.py:12-17


# --------------------------------------------------------------------------- #
# penalties / cost / constraints   (problem.py:38-146)
# --------------------------------------------------------------------------- #
def region_penalty(shapes: List[OShape], X, w: float, smooth: bool, e: float) -> np.ndarray:
    """w * sum_s psi_s(x)/psi_s(center_s)   (problem.py:72-80)."""
    X = np.asarray(X, dtype=F64).reshape(-1, 2)
    total = np.zeros(X.shape[0], dtype=F64)
    with np.errstate(divide='ignore', invalid='ignore'):
        for s in shapes:
            p = s.psi(X, smooth, e)
            if s.center is None or np.isnan(s.center).any():
                total = total + p
            else:
                total = total + p / s.psi(s.center.reshape(1, 2), smooth, e)[0]
    return w * total


def total_penalty(m: OMap, X, weights: Sequence[float], e: float = 0.0, smooth: bool = True) -> np.ndarray:
    """sum over regions in insertion order   (problem.py:49-56)."""
    X = np.asarray(X, dtype=F64).reshape(-1, 2)
    pen = np.zeros(X.shape[0], dtype=F64)
    for (name, shapes), w in zip(m.regions, weights):
        pen = pen + region_penalty(shapes, X, w, smooth, e)
    return pen


def obstacle_penalty(m: OMap, X, e: float = 0.0, smooth: bool = False) -> np.ndarray:
    """get_penalty_function(None): w = 1 over map.obstacles   (problem.py:60-63)."""
    return region_penalty(m.obstacles, X, 1, smooth, e)


def _norm2(D: np.ndarray) -> np.ndarray:
    return np.sqrt(D[..., 0] * D[..., 0] + D[..., 1] * D[..., 1])


def length_of(m: OMap, x, N: int, smooth: bool = False) -> np.ndarray:
    """sum_{k=0}^{N} nrm(y_{k+1}-y_k), y = [x_start; x; x_goal], ONLY the first N+1 pairs
    (problem.py:130-146).  Batched over leading dims of x (…, 2M)."""
    x = np.asarray(x, dtype=F64)
    lead = x.shape[:-1]
    P = x.reshape(lead + (-1, 2))
    ys = np.broadcast_to(m.x_start, lead + (1, 2))
    yg = np.broadcast_to(m.x_goal, lead + (1, 2))
    Y = np.concatenate([ys, P, yg], axis=-2)
    out = np.zeros(lead, dtype=F64)
    for k in range(N + 1):
        d = _norm2(Y[..., k + 1, :] - Y[..., k, :])
        out = out + (d ** 2 if smooth else d)
    return out


def get_cost(m: OMap, z_, N: int, weights: Sequence[float], e: float = 0.0,
             options: Dict = None) -> np.ndarray:
    """(N+1)*length_of(z_) + sum_{j=0}^{N+1} P(z_j)/N   (problem.py:38-44), quirk Q1 included:
    z_ already holds start and goal, so the last segment drops out of the length term."""
    opts = dict(DEFAULT_OPTIONS)
    if options:
        opts.update(options)
    z_ = np.asarray(z_, dtype=F64)
    lead = z_.shape[:-1]
    assert z_.shape[-1] == 2 * (N + 2)
    cost = (N + 1) * length_of(m, z_, N, opts['length_smooth'])
    P = z_.reshape(lead + (N + 2, 2))
    for j in range(N + 2):
        pen = total_penalty(m, P[..., j, :].reshape(-1, 2), weights, e, opts['penalty_smooth'])
        cost = cost + pen.reshape(lead) / N
    return cost


def get_nonlincon(m: OMap, z_, N: int, maxratio: float, maxalpha: float,
                  options: Dict = None) -> np.ndarray:
    """[ratio-hi, ratio-lo, angle] x N  ++  psi_obs(z_j; e=0) per obstacle per waypoint
    (problem.py:84-114; enlargement ignored for obstacles, quirk Q2)."""
    opts = dict(DEFAULT_OPTIONS)
    if options:
        opts.update(options)
    z_ = np.asarray(z_, dtype=F64)
    lead = z_.shape[:-1]
    P = z_.reshape(lead + (N + 2, 2))
    sm = opts['maxratio_smooth']
    nrm = (lambda D: _norm2(D) ** 2) if sm else _norm2
    mr = maxratio ** 2 if sm else maxratio
    mincos = math.cos(maxalpha)
    cols = []
    with np.errstate(divide='ignore', invalid='ignore'):
        for k in range(N):
            zk = P[..., k + 1, :] - P[..., k, :]
            zk1 = P[..., k + 2, :] - P[..., k + 1, :]
            a, b = nrm(zk), nrm(zk1)
            cols.append(np.maximum(0.0, b - mr * a))
            cols.append(np.maximum(0.0, a / mr - b))
            cos_t = (zk[..., 0] * zk1[..., 0] + zk[..., 1] * zk1[..., 1]) / (a * b)
            cols.append(np.maximum(0.0, mincos - cos_t))
    for o in m.obstacles:
        for j in range(N + 2):
            cols.append(o.psi(P[..., j, :].reshape(-1, 2), opts['obstacle_smooth'], 0.0).reshape(lead))
    return np.stack(cols, axis=-1)


def path_collides(m: OMap, z_, N: int) -> np.ndarray:
    """OR over the N+2 waypoints of Map.collides   (map.py:41-43)."""
    z_ = np.asarray(z_, dtype=F64)
    lead = z_.shape[:-1]
    P = z_.reshape(-1, 2)
    return m.collides(P).reshape(lead + (N + 2,)).any(axis=-1)


def create_x_init(x_start, x_goal, N: int, displacement: float = 0.0) -> np.ndarray:
    """Straight line or circular arc through start/goal   (solver.py:103-136)."""
    x0 = np.asarray(x_start, dtype=F64).flatten()
    xf = np.asarray(x_goal, dtype=F64).flatten()
    a = np.linalg.norm(xf - x0) / 2
    if abs(displacement) > 1:
        raise ValueError(f'abs(displacement) = {abs(displacement)} must be smaller than 1')
    out = np.zeros(2 * N)
    if displacement == 0:
        out[0::2] = np.linspace(x0[0], xf[0], N + 2)[1:-1]
        out[1::2] = np.linspace(x0[1], xf[1], N + 2)[1:-1]
        return out
    b = displacement * a
    v = x0 - xf
    alpha = np.arctan2(v[1], v[0])
    R = np.array([[np.cos(alpha), -np.sin(alpha)], [np.sin(alpha), np.cos(alpha)]])
    beta = 2 * np.arctan(2 * a * b / (a ** 2 - b ** 2))
    radius = (a ** 2 + b ** 2) / (2 * b)
    t = np.linspace((np.pi - beta) / 2, (np.pi + beta) / 2, N + 2)[1:-1]
    ell = R @ np.vstack((radius * np.cos(t), (b ** 2 - a ** 2) / (2 * b) + radius * np.sin(t)))
    C = (xf + x0) / 2
    out[0::2] = ell[0, :] + C[0]
    out[1::2] = ell[1, :] + C[1]
    return out


# --------------------------------------------------------------------------- #
# config 1: tests/test_path_generation.py inline problem (:28-66)
# --------------------------------------------------------------------------- #
def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over='ignore'):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def candidate_normals(seed: int, index: np.ndarray, Wp: int) -> np.ndarray:
    """(len(index), Wp, 2) standard normals of the device candidate generator (csrc/uam_candidates.cu: splitmix64 of
    (seed, path index, waypoint) -> two uniforms -> Box-Muller).  Build-defined (the reference jitters nothing)."""
    ctr = (np.asarray(index, dtype=np.uint64)[:, None] * np.uint64(Wp) + np.arange(Wp, dtype=np.uint64)[None, :])
    with np.errstate(over='ignore'):
        h1 = _splitmix64(np.uint64(seed) ^ _splitmix64(np.uint64(2) * ctr))
        h2 = _splitmix64(np.uint64(seed) ^ _splitmix64(np.uint64(2) * ctr + np.uint64(1)))
    u1 = ((h1 >> np.uint64(11)) + np.uint64(1)).astype(F64) * 2.0 ** -53
    u2 = (h2 >> np.uint64(11)).astype(F64) * 2.0 ** -53
    r = np.sqrt(-2.0 * np.log(u1))
    a = 6.283185307179586 * u2
    return np.stack([r * np.cos(a), r * np.sin(a)], axis=-1)


def make_candidates(cand: np.ndarray, N: int, jitter_sigma: float = 0.0, seed: int = 0, index0: int = 0) -> np.ndarray:
    """rows {xs, ys, xg, yg, displacement} -> (B, 2(N+2)) paths [start, create_x_init(displacement), goal]
    (solver.py:103-136 per candidate) + N(0, sigma^2) jitter on the N interior waypoints."""
    cand = np.asarray(cand, dtype=F64).reshape(-1, 5)
    B = cand.shape[0]
    Z = np.empty((B, N + 2, 2), dtype=F64)
    Z[:, 0], Z[:, -1] = cand[:, 0:2], cand[:, 2:4]
    straight = cand[:, 4] == 0.0
    if straight.any():                 # np.linspace on arrays performs the same operations as on scalars (create_x_init, d = 0)
        Z[straight, 1:-1, 0] = np.linspace(cand[straight, 0], cand[straight, 2], N + 2, axis=1)[:, 1:-1]
        Z[straight, 1:-1, 1] = np.linspace(cand[straight, 1], cand[straight, 3], N + 2, axis=1)[:, 1:-1]
    for b in np.nonzero(~straight)[0]:
        Z[b, 1:-1] = create_x_init(cand[b, 0:2], cand[b, 2:4], N, float(cand[b, 4])).reshape(N, 2)
    if jitter_sigma:
        n = candidate_normals(seed, index0 + np.arange(B), N + 2)
        Z[:, 1:-1] += jitter_sigma * n[:, 1:-1]
    return Z.reshape(B, 2 * (N + 2))


def testscript_cost(z, z_start, z_goal, center, R: float = 2.0, w_dist: float = 1.0,
                    w_obs: float = 500.0) -> Tuple[float, float, float]:
    """dist = sum |dz|^2 over N+1 segments; penalty = sum_i max(0, R - |z_i-c|^2)^2 over the N free
    waypoints (R compared with a SQUARED distance, as written at :43-44)."""
    z = np.asarray(z, dtype=F64).reshape(-1, 2)
    pts = np.vstack([np.asarray(z_start, dtype=F64), z, np.asarray(z_goal, dtype=F64)])
    d = np.diff(pts, axis=0)
    dist = float(np.sum(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]))
    q = z - np.asarray(center, dtype=F64)
    d2 = q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]
    pen = float(np.sum(np.maximum(0.0, R - d2) ** 2))
    return dist, pen, w_dist * dist + w_obs * pen


def testscript_constraints(z, z_start, z_goal, r_max: float = 1.1, theta_max: float = math.pi / 6) -> np.ndarray:
    """tests/test_path_generation.py:53-66 (N-1 triples, free waypoints only as pivots)."""
    z = np.asarray(z, dtype=F64).reshape(-1, 2)
    pts = np.vstack([np.asarray(z_start, dtype=F64), z, np.asarray(z_goal, dtype=F64)])
    n = z.shape[0]
    out = []
    for i in range(n - 1):
        dz1 = pts[i + 1] - pts[i]
        dz2 = pts[i + 2] - pts[i + 1]
        a, b = np.linalg.norm(dz1), np.linalg.norm(dz2)
        out += [max(0.0, b - r_max * a), max(0.0, a / r_max - b),
                max(0.0, math.cos(theta_max) - float(np.dot(dz1, dz2)) / (a * b))]
    return np.asarray(out)


# --------------------------------------------------------------------------- #
# raster formulation (BUILD-DEFINED) -- parity unpinned by the reference
# --------------------------------------------------------------------------- #
def cell_centres(n: int, x0: float, dx: float) -> np.ndarray:
    """x0 + (j + 1/2) * dx  -- the exact operation order the device uses."""
    return x0 + (np.arange(n, dtype=F64) + 0.5) * dx


def rasterize_layers(m: OMap, H: int, W: int, x0: float, dx: float, y0: float, dy: float,
                     e: float = 0.0) -> np.ndarray:
    """layer[l, i, j] = float32( sum_{s in region l} psi_s(xc)/psi_s(c_s) ), UNWEIGHTED, smooth."""
    xs = cell_centres(W, x0, dx)
    ys = cell_centres(H, y0, dy)
    out = np.zeros((len(m.regions), H, W), dtype=np.float32)
    for i in range(H):
        X = np.stack([xs, np.full(W, ys[i])], axis=1)
        for l, (name, shapes) in enumerate(m.regions):
            out[l, i] = region_penalty(shapes, X, 1.0, True, e).astype(np.float32)
    return out


def rasterize_occupancy(m: OMap, H: int, W: int, x0: float, dx: float, y0: float, dy: float) -> np.ndarray:
    """occ[i, j] = Map.collides(cell centre)   (map.py:41-43 at cell centres) -- bit-exact target."""
    xs = cell_centres(W, x0, dx)
    ys = cell_centres(H, y0, dy)
    out = np.zeros((H, W), dtype=np.uint8)
    for i in range(H):
        X = np.stack([xs, np.full(W, ys[i])], axis=1)
        out[i] = m.collides(X)
    return out


def rasterize_along_paths(m: OMap, Z: np.ndarray, H: int, W: int, x0: float, dx: float, y0: float, dy: float,
                          samples_per_cell: float = 0.0, e: float = 0.0):
    """rasterize_layers + rasterize_occupancy restricted to the cells score_paths_raster will touch for the paths Z
    (every other cell stays 0): the same texel values at a tiny fraction of the cost, so the CPU tests can tie the raster
    scorer back to the reference's analytic values at 4096^2 and more (a full numpy rasterisation of 4096^2 takes 40 s)."""
    Z = np.asarray(Z, dtype=F64)
    need = np.zeros((H, W), dtype=bool)
    for z in Z:
        P = z.reshape(-1, 2)
        U = (P[:, 0] - x0) / dx - 0.5
        V = (P[:, 1] - y0) / dy - 0.5
        for k in range(len(U) - 1):
            S = int(max(1.0, np.ceil(np.hypot(U[k + 1] - U[k], V[k + 1] - V[k]) * samples_per_cell))) if samples_per_cell > 0 else 1
            s = np.arange(S + 1, dtype=F64)
            u = U[k] + s * ((U[k + 1] - U[k]) / S)
            v = V[k] + s * ((V[k + 1] - V[k]) / S)
            j0 = np.clip(np.floor(u).astype(np.int64), 0, W - 2)
            i0 = np.clip(np.floor(v).astype(np.int64), 0, H - 2)
            for di in (0, 1):
                for dj in (0, 1):
                    need[i0 + di, j0 + dj] = True
    ii, jj = np.nonzero(need)
    X = np.stack([x0 + (jj.astype(F64) + 0.5) * dx, y0 + (ii.astype(F64) + 0.5) * dy], axis=1)
    layers = np.zeros((len(m.regions), H, W), dtype=np.float32)
    occ = np.zeros((H, W), dtype=np.uint8)
    for l, (name, shapes) in enumerate(m.regions):
        layers[l, ii, jj] = region_penalty(shapes, X, 1.0, True, e).astype(np.float32)
    occ[ii, jj] = m.collides(X)
    return layers, occ


def dem_mask(image: np.ndarray, threshold: float = 0.0) -> np.ndarray:
    """image > threshold, or image == -9999 when threshold == -9999   (map_generation/data_manager.py:14-17)."""
    if threshold == -9999:
        return (image == -9999)
    return image > threshold


def pixel_coords(X: np.ndarray, x0: float, dx: float, y0: float, dy: float, H: int, W: int):
    """u = clamp((x-x0)/dx - 1/2, 0, W-1), v likewise; the device repeats this order in fp64."""
    u = (X[..., 0] - x0) / dx - 0.5
    v = (X[..., 1] - y0) / dy - 0.5
    u = np.minimum(np.maximum(u, 0.0), float(W - 1))
    v = np.minimum(np.maximum(v, 0.0), float(H - 1))
    return u, v


def sample_uv(layers: np.ndarray, occ: np.ndarray, u: np.ndarray, v: np.ndarray):
    """Bilinear value per layer (float64 maths on float32 texels) + nearest-cell occupancy, at pixel
    coordinates already clamped to [0,W-1]x[0,H-1].  The fractional parts are rounded to float32
    before use, exactly as the device does (fp64 coordinates -> int cell + fp32 fraction)."""
    L, H, W = layers.shape
    j0 = np.minimum(np.floor(u), W - 2).astype(np.int64)
    i0 = np.minimum(np.floor(v), H - 2).astype(np.int64)
    fx = (u - j0).astype(np.float32).astype(F64)
    fy = (v - i0).astype(np.float32).astype(F64)
    t00 = layers[:, i0, j0].astype(F64)
    t01 = layers[:, i0, j0 + 1].astype(F64)
    t10 = layers[:, i0 + 1, j0].astype(F64)
    t11 = layers[:, i0 + 1, j0 + 1].astype(F64)
    top = t00 + fx * (t01 - t00)
    bot = t10 + fx * (t11 - t10)
    val = top + fy * (bot - top)
    jn = j0 + (fx >= 0.5)
    im = i0 + (fy >= 0.5)
    return val, occ[im, jn] != 0


def score_paths_raster(layers: np.ndarray, occ: np.ndarray, geo: Tuple[float, float, float, float],
                       Z: np.ndarray, weights: Sequence[float], samples_per_cell: float = 0.0,
                       length_smooth: bool = True, x_start=None):
    """Raster path scorer (build-defined).  Z: (B, 2W) interleaved xy incl. start and goal, N = W-2.

    cost = (N+1)*L + (1/N) * [ sum_{k=0}^{N} mean_{s<S_k} P(z_k + s/S_k*(z_{k+1}-z_k)) + P(z_{N+1}) ]
    with P = sum_l w_l * bilinear(layer_l), L = the reference's length term incl. quirk Q1
    (problem.py:38-44,130-146; `x_start` = map.x_start, None -> each path's own z_0 so the term is 0),
    S_k = 1 when samples_per_cell == 0 (waypoint mode == the reference's sampling) else
    max(1, ceil(|dz_k|_pixels * samples_per_cell)).  collide = OR over samples of nearest-cell occupancy.
    Returns (cost float64 (B,), collide bool (B,), n_samples int64 (B,)).
    """
    x0, dx, y0, dy = geo
    L_, H, W_ = layers.shape
    Z = np.asarray(Z, dtype=F64)
    B = Z.shape[0]
    Wp = Z.shape[1] // 2
    N = Wp - 2
    P = Z.reshape(B, Wp, 2)
    w = np.asarray(weights, dtype=F64)
    # length term (quirk Q1): |z_0 - m_s| + sum_{k=0}^{N-1} |dz_k|
    D = P[:, 1:N + 1, :] - P[:, 0:N, :]
    d = _norm2(D)
    Lsum = np.sum(d ** 2 if length_smooth else d, axis=1)
    if x_start is not None:
        d0 = _norm2(P[:, 0, :] - np.asarray(x_start, dtype=F64))
        Lsum = Lsum + (d0 ** 2 if length_smooth else d0)
    # unclamped pixel coordinates of the waypoints (fp64), clamped per sample
    U = (P[..., 0] - x0) / dx - 0.5
    V = (P[..., 1] - y0) / dy - 0.5
    dU = U[:, 1:] - U[:, :-1]
    dV = V[:, 1:] - V[:, :-1]
    if samples_per_cell > 0:
        S = np.maximum(1.0, np.ceil(np.sqrt(dU * dU + dV * dV) * samples_per_cell)).astype(np.int64)
    else:
        S = np.ones((B, Wp - 1), dtype=np.int64)
    pen = np.zeros(B, dtype=F64)
    col = np.zeros(B, dtype=bool)
    nsmp = S.sum(axis=1) + 1
    Smax = int(S.max())
    for k in range(Wp - 1):
        Sk = S[:, k]
        su, sv = dU[:, k] / Sk, dV[:, k] / Sk
        acc = np.zeros(B, dtype=F64)
        for s in range(int(Sk.max())):
            act = s < Sk
            u = np.minimum(np.maximum(U[:, k] + s * su, 0.0), float(W_ - 1))
            v = np.minimum(np.maximum(V[:, k] + s * sv, 0.0), float(H - 1))
            val, o = sample_uv(layers, occ, u, v)
            acc += np.where(act, w @ val, 0.0)
            col |= (o & act)
        pen += acc / Sk
    u = np.minimum(np.maximum(U[:, -1], 0.0), float(W_ - 1))
    v = np.minimum(np.maximum(V[:, -1], 0.0), float(H - 1))
    val, o = sample_uv(layers, occ, u, v)
    pen += w @ val
    col |= o
    cost = (N + 1) * Lsum + pen / N
    return cost, col, nsmp


# --------------------------------------------------------------------------- #
# map rebuild extensions: exact EDT, grid search  (no reference counterpart; parity unpinned)
# --------------------------------------------------------------------------- #
def edt_sq(occ: np.ndarray) -> np.ndarray:
    """Exact squared Euclidean distance (in cells, int64) from every cell to the nearest occupied
    cell; occupied cells get 0; a map with no occupied cell gets a large sentinel (2**30)."""
    from scipy import ndimage
    if not occ.any():
        return np.full(occ.shape, 2 ** 30, dtype=np.int64)
    d = ndimage.distance_transform_edt(occ == 0)
    return np.rint(d * d).astype(np.int64)


def grid_search(cost: np.ndarray, start, blocked: np.ndarray = None):
    """Cost-to-go on an 8-connected grid with integer edge costs (Dijkstra, exact), optionally with altitude bands.

    cost (H,W) + start (row, col), or cost (bands,H,W) + start (band, row, col).
    In-plane edge u->v costs step(u,v) * (cost[b,u] + cost[b,v]) with step = 2 for axis moves and 3 for diagonal
    moves (integer 2:3 approximation of 1:sqrt2); a band change at a fixed cell costs 2 * (cost[b,v] + cost[b+-1,v]).
    All integer, so results are order-independent.  dist[start] = 0; unreachable / blocked = 2**62.
    parent[v] = flat index of the neighbour u that minimises dist[u] + w(u,v), ties -> smallest neighbour slot in the
    fixed order (-1,-1),(-1,0),(-1,1),(0,-1),(0,1),(1,-1),(1,0),(1,1), band below, band above; parent[start] = start;
    unreachable = -1.
    """
    import heapq
    flat2d = cost.ndim == 2
    if flat2d:
        cost = cost[None]
        blocked = None if blocked is None else blocked[None]
        start = (0,) + tuple(start)
    Bn, H, W = cost.shape
    cells = H * W
    INF = 2 ** 62
    dist = np.full(Bn * cells, INF, dtype=np.int64)
    c = cost.astype(np.int64).ravel().tolist()
    blk = (np.zeros(Bn * cells, dtype=bool) if blocked is None else (blocked.ravel() != 0)).tolist()
    s = start[0] * cells + start[1] * W + start[2]
    nb = [(-1, -1, 3), (-1, 0, 2), (-1, 1, 3), (0, -1, 2), (0, 1, 2), (1, -1, 3), (1, 0, 2), (1, 1, 3)]

    def neighbours(v):
        vb, vc = divmod(v, cells)
        vi, vj = divmod(vc, W)
        for di, dj, st in nb:
            ui, uj = vi + di, vj + dj
            if 0 <= ui < H and 0 <= uj < W:
                yield vb * cells + ui * W + uj, st
        for db in (-1, 1):
            if 0 <= vb + db < Bn:
                yield (vb + db) * cells + vc, 2

    dl = dist.tolist()
    if not blk[s]:
        dl[s] = 0
        pq = [(0, s)]
        while pq:
            d, u = heapq.heappop(pq)
            if d != dl[u]:
                continue
            for v, st in neighbours(u):
                if blk[v]:
                    continue
                nd = d + st * (c[u] + c[v])
                if nd < dl[v]:
                    dl[v] = nd
                    heapq.heappush(pq, (nd, v))
    parent = [-1] * (Bn * cells)
    for v in range(Bn * cells):
        if dl[v] >= INF:
            continue
        if v == s:
            parent[v] = s
            continue
        best, bp = INF, -1
        for u, st in neighbours(v):
            if dl[u] >= INF:
                continue
            nd = dl[u] + st * (c[u] + c[v])
            if nd < best:
                best, bp = nd, u
        parent[v] = bp
    shape = (H, W) if flat2d else (Bn, H, W)
    return np.array(dl, dtype=np.int64).reshape(shape), np.array(parent, dtype=np.int64).reshape(shape)


# --------------------------------------------------------------------------- #
# gradient of get_cost w.r.t. the waypoints (extension: the reference gets it from CasADi's AD inside OpEn,
# solver.py:82-101; no in-tree counterpart).  Pinned by central finite differences of get_cost above.
# --------------------------------------------------------------------------- #
def _grad_h(edge, X):
    x0, x1 = X[:, 0], X[:, 1]
    if edge[0] == 'line':
        _, Ax, Ay, Bx, By, sgn = edge
        return np.stack([np.full_like(x0, -sgn * (By - Ay)), np.full_like(x0, sgn * (Bx - Ax))], axis=1)
    if edge[0] == 'ellipse':
        _, cx, cy, r1, r2 = edge
        return np.stack([2 * ((x0 - cx) / r1) / r1, 2 * ((x1 - cy) / r2) / r2], axis=1)
    _, axis, sign, c, r = edge
    g = np.zeros((X.shape[0], 2))
    g[:, axis] = 1.0 if sign > 0 else -1.0
    return g


def psi_grad(shape: OShape, X, e: float = 0.0):
    """(psi, grad psi) at points X for the smooth penalty: grad psi = psi * sum_i 2 grad h_i / m_i where all m_i < 0."""
    X = np.asarray(X, dtype=F64).reshape(-1, 2)
    Hm = np.minimum(shape.h(X) - e, 0.0)                       # (n_edges, M)
    psi = np.prod(Hm ** 2, axis=0) if Hm.shape[0] else np.ones(X.shape[0])
    inside = np.all(Hm < 0, axis=0)
    s = np.zeros((X.shape[0], 2))
    with np.errstate(divide='ignore', invalid='ignore'):
        for i, edge in enumerate(shape.edges):
            s += np.where(inside[:, None], 2 * _grad_h(edge, X) / Hm[i][:, None], 0.0)
    return psi, psi[:, None] * s


def get_cost_gradient(m: OMap, z_, N: int, weights: Sequence[float], e: float = 0.0, options: Dict = None) -> np.ndarray:
    """d get_cost / d z_ (same shape as z_, start and goal columns included), penalty_smooth only."""
    opts = dict(DEFAULT_OPTIONS)
    if options:
        opts.update(options)
    assert opts['penalty_smooth']
    z_ = np.asarray(z_, dtype=F64)
    B = z_.shape[0]
    P = z_.reshape(B, N + 2, 2)
    G = np.zeros_like(P)
    for (name, shapes), w in zip(m.regions, weights):
        for s in shapes:
            psi, gp = psi_grad(s, P.reshape(-1, 2), e)
            pc = s.psi(s.center.reshape(1, 2), True, e)[0] if (s.center is not None and not np.isnan(s.center).any()) else 1.0
            G += (w * gp / pc).reshape(B, N + 2, 2) / N
    # length term: pairs (m_s, z_0), (z_0, z_1), ..., (z_{N-1}, z_N)
    Y = np.concatenate([np.broadcast_to(m.x_start, (B, 1, 2)), P], axis=1)      # Y_0 = m_s, Y_{k+1} = z_k
    for k in range(N + 1):
        d = Y[:, k + 1] - Y[:, k]
        if opts['length_smooth']:
            gk = 2 * d
        else:
            n = _norm2(d)[:, None]
            gk = np.divide(d, n, out=np.zeros_like(d), where=n > 0)      # coincident points: the zero subgradient
        G[:, k] += (N + 1) * gk                   # d/d z_k   (z_k = Y_{k+1})
        if k >= 1:
            G[:, k - 1] -= (N + 1) * gk           # d/d z_{k-1}
    return G.reshape(B, 2 * (N + 2))


# ---------------------------------------------------------------------------------------------------------------
# polygon front-end of map_generation: mask -> connected regions -> minimum-area rectangles
# ---------------------------------------------------------------------------------------------------------------
def label_components(mask, connectivity: int = 4):
    """One label per connected region of the mask, numbered in raster-scan order of the region's first cell
    (scipy.ndimage.label).  The reference gets the same regions as polygons from rasterio.features.shapes
    (map_generation/data_manager.py:18-19; 4-connectivity is rasterio's default)."""
    from scipy import ndimage
    st = ndimage.generate_binary_structure(2, 1 if connectivity == 4 else 2)
    lab, n = ndimage.label(np.asarray(mask) != 0, structure=st)
    return lab.astype(np.int32), int(n)


def component_stats(labels, n: int):
    """cells per component (= polygon.area / cell area, map_generation/data_processor.py:19) and bounding boxes
    {row min, row max, col min, col max}."""
    area = np.bincount(labels.ravel(), minlength=n + 1)[1:].astype(np.int64)
    bbox = np.zeros((n, 4), dtype=np.int32)
    ii, jj = np.nonzero(labels)
    lab = labels[ii, jj] - 1
    for k, (arr, fn) in enumerate([(ii, np.minimum), (ii, np.maximum), (jj, np.minimum), (jj, np.maximum)]):
        v = np.full(n, np.iinfo(np.int32).max if fn is np.minimum else -1, dtype=np.int64)
        fn.at(v, lab, arr)
        bbox[:, k] = v
    return area, bbox


def _hull_int(points):
    """Strictly convex hull of integer points (Andrew's monotone chain, exact), as a list of (x, y) tuples."""
    pts = sorted(set((int(x), int(y)) for x, y in points))
    if len(pts) <= 2:
        return pts

    def half(seq):
        h = []
        for p in seq:
            while len(h) >= 2 and (h[-1][0] - h[-2][0]) * (p[1] - h[-2][1]) - (h[-1][1] - h[-2][1]) * (p[0] - h[-2][0]) <= 0:
                h.pop()
            h.append(p)
        return h
    lo, up = half(pts), half(pts[::-1])
    return lo[:-1] + up[:-1]


def min_area_rect_exact(points):
    """Minimum-area enclosing rectangle of integer points: a side of it lies on a hull edge, so every hull edge is
    tried; extents are exact integers, areas are compared as exact fractions, ties go to the first edge of the hull
    listed from its top-most (then left-most) vertex down the left side (y grows downwards, raster rows).
    -> (corners (4,2) float64, area as a Fraction, number of hull vertices, index of the chosen edge).
    This is what cv2.minAreaRect approximates in float32 (map_generation/data_processor.py:67-71)."""
    from fractions import Fraction
    h = _hull_int(points)
    n = len(h)
    assert n >= 3
    # orientation: cross(b - a, c - a) < 0 for consecutive vertices (down the left side first when y grows downwards)
    a, b, c = h[0], h[1], h[2]
    if (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]) > 0:
        h = h[::-1]
    s = min(range(n), key=lambda i: (h[i][1], h[i][0]))
    h = h[s:] + h[:s]
    best = None
    for e in range(n):
        ax, ay = h[e]
        bx, by = h[(e + 1) % n]
        ex, ey = bx - ax, by - ay
        us = [(px - ax) * ex + (py - ay) * ey for px, py in h]
        vs = [abs((px - ax) * ey - (py - ay) * ex) for px, py in h]
        den = ex * ex + ey * ey
        area = Fraction((max(us) - min(us)) * max(vs), den)
        if best is None or area < best[0]:
            best = (area, e, min(us), max(us), max(vs), den, (ax, ay, ex, ey))
    area, e, u0, u1, vv, den, (ax, ay, ex, ey) = best
    nx, ny = ey, -ex
    u0, u1, vv = u0 / den, u1 / den, vv / den
    corners = np.array([[ax + u0 * ex, ay + u0 * ey], [ax + u1 * ex, ay + u1 * ey],
                        [ax + u1 * ex + vv * nx, ay + u1 * ey + vv * ny], [ax + u0 * ex + vv * nx, ay + u0 * ey + vv * ny]], dtype=F64)
    return corners, area, n, e


def component_rects(labels, ids, geo):
    """Minimum-area rectangle of the cell corners of each listed component, world coordinates (x0 + col dx, y0 + row dy)."""
    x0, dx, y0, dy = geo
    out = np.zeros((len(ids), 4, 2), dtype=F64)
    info = np.zeros((len(ids), 2), dtype=np.int32)
    for k, lab in enumerate(ids):
        ii, jj = np.nonzero(labels == lab)
        pts = np.concatenate([np.stack([jj + a, ii + b], 1) for a in (0, 1) for b in (0, 1)])
        c, _, nh, e = min_area_rect_exact(pts)
        out[k, :, 0] = x0 + c[:, 0] * dx
        out[k, :, 1] = y0 + c[:, 1] * dy
        info[k] = (nh, e)
    return out, info


def split_component_rects(labels, label: int, bbox, geo, divisions: int = 5):
    """DataProcessor._divide_and_approximate_polygon (map_generation/data_processor.py:34-53) restated on the raster: the
    component's bounding box is cut into divisions x divisions boxes; on the grid refined `divisions` times a box is a block
    of sub-cells, its 4-connected regions are the pieces of polygon.intersection(box), and each piece gets the exact
    minimum-area rectangle of its sub-cell corners.  Box order as in the reference (x index outer, y index inner, counted
    from minx / miny).  -> (rects (P,4,2) float64 world coordinates, box index j * divisions + k per piece)."""
    x0, dx, y0, dy = [float(v) for v in geo]
    r0, r1, c0, c1 = [int(v) for v in bbox]
    nr, nc = r1 - r0 + 1, c1 - c0 + 1
    fine = np.kron((labels[r0:r1 + 1, c0:c1 + 1] == label).astype(np.uint8), np.ones((divisions, divisions), dtype=np.uint8))
    rows = list(range(divisions)) if dy > 0 else list(range(divisions - 1, -1, -1))
    cols = list(range(divisions)) if dx > 0 else list(range(divisions - 1, -1, -1))
    rects, boxes = [], []
    for j, bc in enumerate(cols):
        for k, br in enumerate(rows):
            sub = fine[br * nr:(br + 1) * nr, bc * nc:(bc + 1) * nc]
            lab, n = label_components(sub, 4)
            gx0, gy0 = x0 + (c0 + bc * nc / divisions) * dx, y0 + (r0 + br * nr / divisions) * dy
            for p in range(1, n + 1):
                ii, jj = np.nonzero(lab == p)
                pts = np.concatenate([np.stack([jj + a, ii + b], axis=1) for a in (0, 1) for b in (0, 1)])
                corners, _, _, _ = min_area_rect_exact(pts)
                rects.append(np.stack([gx0 + corners[:, 0] * (dx / divisions), gy0 + corners[:, 1] * (dy / divisions)], axis=1))
                boxes.append(j * divisions + k)
    return (np.stack(rects) if rects else np.zeros((0, 4, 2))), np.asarray(boxes, dtype=np.int32)


def grid_path(parent: np.ndarray, start, goal):
    """Node list (flat indices into the grid) from start to goal along the predecessors of ``grid_search``; empty when the
    goal was not reached.  start / goal: (row, col) or (band, row, col)."""
    shape = parent.shape
    if len(shape) == 2:
        start, goal = (0,) + tuple(start), (0,) + tuple(goal)
        shape = (1,) + shape
    p = parent.reshape(-1)
    cells = shape[1] * shape[2]
    s = start[0] * cells + start[1] * shape[2] + start[2]
    v = goal[0] * cells + goal[1] * shape[2] + goal[2]
    out = [v]
    while v != s:
        v = int(p[v])
        if v < 0:
            return []
        out.append(v)
    return out[::-1]

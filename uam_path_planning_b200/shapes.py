"""Shape constructors of the planner's map: host-side model build.

Mirrors ``polygon(*points)``, ``ball(center, r1, r2)``, ``square(center, r1, r2)`` and the
``QuadraticObstacle`` they return (reference: path_generation/polygon.py:7-143, ball.py:7-52,
square.py:6-65, quadratic_obstacle.py:8-39,89-94).  A shape here is not a list of Python closures but a
small table of inequality records -- the rows ``uam_map_set_shapes`` uploads (include/uam_b200.h):

    line    [0, Ax, Ay, Bx-Ax, By-Ay, sgn, 0, 0]   h = -sgn*((By-Ay)(x-Ax) - (Bx-Ax)(y-Ay))
    ellipse [1, cx, cy, r1, r2, 0, 0, 0]           h = ((x-cx)/r1)^2 + ((y-cy)/r2)^2 - 1
    box     [2, axis, sign, c, r, 0, 0, 0]         h = sign>0 ? x_d - c - r : -x_d + c - r

Only the table is built on the host; every evaluation (``contains``, ``penalty_function``) runs on the GPU.
"""
from __future__ import annotations

import itertools
import math
from typing import List, Optional, Sequence

import numpy as np

from ._lib import UAM_EDGE_BOX, UAM_EDGE_ELLIPSE, UAM_EDGE_LINE


_SERIAL = itertools.count(1)      # every Inequality / shape state gets a fresh number: cheap, collision-free change detection
EPOCH = [0]                       # bumped by every such change anywhere: "nothing changed since" is one integer compare


class Inequality:
    """One record h(x) <= 0 of a shape (stands in for the reference's ``Function``, function.py:4-120).  The record array
    is read-only: an edit means assigning a new record (``ineq.record = ...``), which the maps' device tables notice."""
    __slots__ = ('_record', '_serial', 'is_quadratic', 'n')

    def __init__(self, record):
        self.record = record
        self.is_quadratic = True
        self.n = 2

    @property
    def record(self) -> np.ndarray:
        return self._record

    @record.setter
    def record(self, value):
        r = np.array(value, dtype=np.float64).reshape(8)
        r.flags.writeable = False
        self._record = r
        self._serial = next(_SERIAL)
        EPOCH[0] += 1

    def __call__(self, x):
        from .engine import default_engine
        out = default_engine().eval_inequalities([self.record], x)[0]
        return float(out[0]) if np.ndim(x) == 1 else out


class QuadraticObstacle:
    """A convex shape = ordered inequality records + ``center`` + ``area`` (quadratic_obstacle.py:8-39)."""

    def __init__(self, *inequalities: Inequality):
        self.inequalities: List[Inequality] = []
        self.area = float('nan')
        self._center = float('nan')
        self._serial = next(_SERIAL)
        self.xlim: Optional[List[float]] = None
        self.ylim: Optional[List[float]] = None
        self.kind = 'generic'
        self.add(*inequalities)

    def add(self, *inequalities):
        for ineq in inequalities:
            assert isinstance(ineq, Inequality), f'Expected Inequality, got {type(ineq)}'
            assert ineq.n == 2, f'Function must be 2-dimensional, got {ineq.n}-dimensional'
            self.inequalities.append(ineq)
            EPOCH[0] += 1

    @property
    def center(self):
        """``QuadraticObstacle.center`` (assignable like the reference's attribute; array values are stored read-only, so
        a change is always an assignment and bumps the shape's state number)."""
        return self._center

    @center.setter
    def center(self, value):
        if isinstance(value, np.ndarray):
            value = value.copy()
            value.flags.writeable = False
        self._center = value
        self._serial = next(_SERIAL)
        EPOCH[0] += 1

    def state_key(self):
        """Changes whenever something the device table depends on changes (centre, inequality list, any record)."""
        return (self._serial, tuple(h._serial for h in self.inequalities))

    def __len__(self):
        """Number of inequalities (quadratic_obstacle.py:211-213)."""
        return len(self.inequalities)

    # ---- table form -------------------------------------------------------------------------------
    def records(self) -> np.ndarray:
        if not self.inequalities:
            return np.zeros((0, 8), dtype=np.float64)
        return np.stack([h.record for h in self.inequalities])

    def center_or_nan(self) -> np.ndarray:
        c = np.asarray(self.center, dtype=np.float64)
        if c.shape != (2,):
            return np.array([np.nan, np.nan])
        return c

    # ---- device-evaluated queries ---------------------------------------------------------------------
    def contains(self, x) -> bool:
        """all_i h_i(x) <= 1e-14   (quadratic_obstacle.py:89-94); (M,2) input gives a bool array."""
        from .engine import default_engine
        out = default_engine().eval_single_shape(self, x, want='contains')
        return bool(out[0]) if np.ndim(x) == 1 else out

    def penalty_function(self, smooth=True, enlargement=0):
        """psi(x) = prod_i min(h_i(x)-e, 0)^2 | prod_i min(e-h_i(x), 0)   (quadratic_obstacle.py:27-39)."""
        def psi(x):
            from .engine import default_engine
            out = default_engine().eval_single_shape(self, x, want='psi', smooth=smooth, enlargement=enlargement)
            return float(out[0]) if np.ndim(x) == 1 else out
        return psi


def _point(p) -> np.ndarray:
    return np.array(p).reshape(2, 1)


def polygon(*points) -> QuadraticObstacle:
    """Convex polygon through P1, P2, ... (any order): gift-wrap from vertex 0 with a convexity check
    (polygon.py:20-21,55-136).  Edge k joins consecutive hull vertices A -> B; its sign is the side of the
    line on which every other vertex lies.  Error messages are the reference's."""
    if len(points) < 3:
        raise ValueError(f'Only {len(points)} vertices given. At least 3 required')
    pts = [_point(p) for p in points]
    n = len(pts)
    # center: running sum in the dtype of the first vertex, exactly like `center = Pa.copy(); center += Pb`
    # (an integer first vertex followed by float vertices raises numpy's casting error, as in the reference)
    center = pts[0].copy()
    for b in range(1, n):
        center += pts[b]
    V = np.concatenate(pts, axis=1).astype(np.float64)      # (2, n)
    vx, vy = V[0], V[1]

    def side_signs(a, b):
        """sign of (By-Ay)(x-Ax) - (Bx-Ax)(y-Ay) at every other vertex (polygon.py:69-71)."""
        line = (vy[b] - vy[a]) * (vx - vx[a]) - (vx[b] - vx[a]) * (vy - vy[a])
        keep = np.ones(n, dtype=bool)
        keep[a] = keep[b] = False
        return np.sign(line[keep])

    def consecutive(a, b):
        sg = side_signs(a, b)
        first = 0.0
        for s in sg:                     # the reference raises/returns at the first offending vertex
            if s == 0:
                raise ValueError('Input contains three aligned points')
            if first == 0:
                first = s
            elif s != first:
                return None
        if first == 0:
            raise ValueError('The polygon is nonconvex')
        return [UAM_EDGE_LINE, vx[a], vy[a], vx[b] - vx[a], vy[b] - vy[a], float(first), 0.0, 0.0]

    obs = QuadraticObstacle()
    obs.kind = 'polygon'
    remaining = list(range(1, n))
    a = 0
    area = 0.0
    while remaining:
        for idx, b in enumerate(remaining):
            rec = consecutive(a, b)
            if rec is not None:
                area += vx[a] * vy[b] - vy[a] * vx[b]
                remaining.pop(idx)
                a = b
                obs.add(Inequality(rec))
                break
        else:
            raise ValueError('The polygon is nonconvex')
    rec = consecutive(a, 0)
    if rec is None:
        raise ValueError("Couldn't close polygon")
    area += vx[a] * vy[0] - vy[a] * vx[0]
    obs.add(Inequality(rec))
    obs.xlim = [float(vx.min()), float(vx.max())]
    obs.ylim = [float(vy.min()), float(vy.max())]
    obs.area = abs(area) / 2
    obs.center = (center / n).reshape(2)
    return obs


def ball(center, r1: float = None, r2: float = None) -> QuadraticObstacle:
    """Ellipse of radii r1, r2 at `center`; a single argument is the radius of a disc at the origin
    (ball.py:19-24,33-37,49-50)."""
    if r1 is None and r2 is None:
        r1 = center
        r2 = r1
        center = np.array([0.0, 0.0])
    elif r2 is None:
        r2 = r1
    center = np.array(center)
    assert center.shape == (2,)
    obs = QuadraticObstacle(Inequality([UAM_EDGE_ELLIPSE, center[0], center[1], r1, r2, 0.0, 0.0, 0.0]))
    obs.kind = 'ball'
    obs.xlim = [float(center[0] - r1), float(center[0] + r1)]
    obs.ylim = [float(center[1] - r2), float(center[1] + r2)]
    obs.center = center
    obs.area = math.pi * r1 * r2
    return obs


def square(center, r1: float, r2: Optional[float] = None) -> QuadraticObstacle:
    """Axis-aligned box with half sides r1, r2; sides in the order right, left, top, bottom (square.py:18-65)."""
    center = np.array(center).reshape(2)
    if r2 is None:
        r2 = r1
    recs = [[UAM_EDGE_BOX, 0, +1, center[0], r1, 0, 0, 0], [UAM_EDGE_BOX, 0, -1, center[0], r1, 0, 0, 0],
            [UAM_EDGE_BOX, 1, +1, center[1], r2, 0, 0, 0], [UAM_EDGE_BOX, 1, -1, center[1], r2, 0, 0, 0]]
    obs = QuadraticObstacle(*[Inequality(r) for r in recs])
    obs.kind = 'square'
    obs.xlim = [float(center[0] - r1), float(center[0] + r1)]
    obs.ylim = [float(center[1] - r2), float(center[1] + r2)]
    obs.center = center
    obs.area = 4 * r1 * r2
    return obs


def flatten_shapes(obstacles: Sequence[QuadraticObstacle], regions: Sequence[Sequence[QuadraticObstacle]]):
    """Shape lists -> the arrays of ``uam_map_set_shapes``: (edges (E,8) f64, off (S+1) i32, region (S) i32,
    center (S,2) f64).  Obstacles get region -1, region shapes their region's insertion index."""
    recs, off, reg, cen = [], [0], [], []
    for r, shapes in [(-1, obstacles)] + list(enumerate(regions)):
        for s in shapes:
            R = s.records()
            recs.append(R)
            off.append(off[-1] + R.shape[0])
            reg.append(r)
            cen.append(s.center_or_nan())
    edges = np.concatenate(recs, axis=0) if recs else np.zeros((0, 8))
    return (np.ascontiguousarray(edges, dtype=np.float64), np.asarray(off, dtype=np.int32),
            np.asarray(reg, dtype=np.int32), np.ascontiguousarray(np.asarray(cen, dtype=np.float64).reshape(-1, 2)))

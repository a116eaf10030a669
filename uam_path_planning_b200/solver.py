"""``Solver``: candidate generation and best-of-candidates selection around the GPU scorer.

Keeps ``Solver(problem, opts)`` and ``create_x_init(displacement)`` of path_generation/solver.py:8-17,103-136
and the best-candidate bookkeeping of path_generation/main.py:160-180.  ``solve`` / ``build_solver`` drive the
OpEn NLP solver (opengen + cargo + TCP) in the reference; that optimiser is not part of the accelerated path
and is not reimplemented.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from .problem import Problem


class Solver:
    def __init__(self, problem: Problem, opts: Dict = None):
        assert isinstance(problem, Problem)
        self.problem = problem
        self.x_sol = None
        self.x_init = None
        self.opts = opts
        self.verbose = True
        self.optimizer_name = None
        self.update_solver = False

    def solve(self, x_init, params):
        raise NotImplementedError('the OpEn/PANOC optimiser of the reference (solver.py:19-101) is out of scope; '
                                  'use evaluate_candidates() to score and rank candidate paths on the GPU')

    build_solver = solve

    def get_error_code_explanation(self, error_code):
        """Texts of the OpEn TCP server's error codes (solver.py:169-177), kept for callers that report them."""
        return {1000: 'Invalid request: Malformed or invalid JSON', 1600: 'Initial guess has incompatible dimensions',
                1700: 'Wrong dimension of Langrange multipliers', 2000: 'Problem solution failed (solver error)',
                3003: 'Vector `parameter` has wrong length'}.get(error_code, 'Error code not found')

    def create_x_init(self, displacement=0):
        """Straight line (displacement 0) or the circular arc through start and goal whose sagitta is
        displacement * |goal - start| / 2; N interior points equally spaced in angle (solver.py:103-136)."""
        N = self.problem.N
        x0 = np.array(self.problem.map.x_start).flatten()
        xf = np.array(self.problem.map.x_goal).flatten()
        a = np.linalg.norm(xf - x0) / 2
        if abs(displacement) > 1:
            raise ValueError(f'abs(displacement) = {abs(displacement)} must be smaller than 1')
        out = np.zeros(2 * N)
        if displacement == 0:
            out[0::2] = np.linspace(x0[0], xf[0], N + 2)[1:-1]
            out[1::2] = np.linspace(x0[1], xf[1], N + 2)[1:-1]
            return out
        b = displacement * a                       # distance of the chord from the arc's apex
        v = x0 - xf
        alpha = np.arctan2(v[1], v[0])
        rot = np.array([[np.cos(alpha), -np.sin(alpha)], [np.sin(alpha), np.cos(alpha)]])
        beta = 2 * np.arctan(2 * a * b / (a ** 2 - b ** 2))
        radius = (a ** 2 + b ** 2) / (2 * b)
        t = np.linspace((np.pi - beta) / 2, (np.pi + beta) / 2, N + 2)[1:-1]
        arc = rot @ np.vstack((radius * np.cos(t), (b ** 2 - a ** 2) / (2 * b) + radius * np.sin(t)))
        mid = (xf + x0) / 2
        out[0::2] = arc[0, :] + mid[0]
        out[1::2] = arc[1, :] + mid[1]
        return out

    # ---- batched forms ----------------------------------------------------------------------------------
    def full_path(self, x) -> np.ndarray:
        """[x_start, x, x_goal] per row: the z_ layout of solver.py:64-66."""
        X = np.atleast_2d(np.asarray(x, dtype=np.float64))
        s = np.broadcast_to(np.asarray(self.problem.map.x_start, dtype=np.float64).ravel(), (X.shape[0], 2))
        g = np.broadcast_to(np.asarray(self.problem.map.x_goal, dtype=np.float64).ravel(), (X.shape[0], 2))
        return np.ascontiguousarray(np.concatenate([s, X, g], axis=1))

    def candidates(self, displacements: Sequence[float], jitter: float = 0.0,
                   rng: Optional[np.random.Generator] = None) -> np.ndarray:
        """(B, 2(N+2)) candidate paths: one arc per displacement, optional Gaussian jitter (same units as the
        map) on the interior waypoints -- the arc family of main.py:160-170 swept densely."""
        X = np.stack([self.create_x_init(float(d)) for d in displacements])
        if jitter:
            rng = rng or np.random.default_rng(0)
            X = X + rng.normal(0.0, jitter, X.shape)
        return self.full_path(X)

    def candidates_device(self, displacements):
        """The same arc family generated on the GPU from a CUDA tensor of displacements (8 bytes of input per
        candidate): (B, 2(N+2)) float64 CUDA tensor, ready for Problem.score / RasterMap.score_paths."""
        m = self.problem.map
        return m.engine().make_arc_paths(m.x_start, m.x_goal, self.problem.N, displacements)

    def seed_from_grid_path(self, nodes, raster_shape, geo) -> np.ndarray:
        """A grid-search path (flat node ids of Engine.grid_paths, source -> goal, on an (H,W) or (bands,H,W) grid whose
        cell centres are (x0 + (j + 1/2) dx, y0 + (i + 1/2) dy)) resampled at N points equally spaced in arc length:
        the flat (2N,) vector Solver.create_x_init returns (solver.py:103-136), i.e. a seed for the scorer / optimiser
        that follows the cheapest grid route instead of a circular arc.  The path's ends are replaced by map.x_start /
        map.x_goal (they lie within a cell of them when the query was made from those points)."""
        H, W = raster_shape[-2:]
        x0, dx, y0, dy = [float(v) for v in geo]
        v = np.asarray(nodes, dtype=np.int64).reshape(-1) % (H * W)
        pts = np.stack([x0 + (v % W + 0.5) * dx, y0 + (v // W + 0.5) * dy], axis=1)
        m = self.problem.map
        pts = np.concatenate([[np.asarray(m.x_start, dtype=np.float64)], pts[1:-1], [np.asarray(m.x_goal, dtype=np.float64)]])
        seg = np.sqrt(((pts[1:] - pts[:-1]) ** 2).sum(axis=1))
        keep = np.concatenate([[True], seg > 0])             # a band change repeats a cell
        pts, seg = pts[keep], seg[seg > 0]
        s = np.concatenate([[0.0], np.cumsum(seg)])
        N = self.problem.N
        t = np.linspace(0.0, s[-1], N + 2)[1:-1]
        return np.stack([np.interp(t, s, pts[:, 0]), np.interp(t, s, pts[:, 1])], axis=1).reshape(-1)

    def evaluate_candidates(self, Z) -> Dict:
        """Score a batch and pick the best like main.py:162-180: fval = sqrt(cost) (solver.py:48), strict `<`
        so ties keep the earliest candidate; also the shortest by the non-smooth length (solver.py:49)."""
        prob = self.problem
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        cost, collide, _ = prob.score(Z)
        fval = np.sqrt(cost)
        length = prob.length_of(np.ascontiguousarray(Z[:, 2:-2]), False)
        best = int(np.argmin(np.where(np.isnan(fval), np.inf, fval)))
        return {'cost': cost, 'fval': fval, 'collide': collide.astype(bool), 'length': length,
                'min_fval': float(fval[best]), 'min_fval_index': best,
                'min_length': float(np.min(length)), 'min_length_index': int(np.argmin(length))}

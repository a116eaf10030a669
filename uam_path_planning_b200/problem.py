"""``Problem``: cost, penalties, constraint vector and path length of candidate paths, evaluated on the GPU.

Drop-in for path_generation/problem.py:6-146 -- same constructor, ``options`` / ``params`` / ``weights``
dictionaries and method names -- with every method also accepting a batch: ``z`` may be one flat path
``[xs,ys,x1,y1,...,xN,yN,xg,yg]`` (returns a scalar / vector like the reference) or a ``(B, 2(N+2))`` array or
CUDA tensor (returns ``(B,)`` / ``(B, len(g))``).  The reference's quirks are kept: the
length term of ``get_cost`` drops the last segment, obstacles in ``get_nonlincon`` ignore ``enlargement``,
non-smooth penalties give NaN costs.
"""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np

from . import _lib
from .engine import _is_tensor
from .region_map import RegionMap


class Problem:
    def __init__(self, map: RegionMap, N: int, opts: Dict = None):
        assert isinstance(map, RegionMap)
        self.map = map
        self.N = N
        self.weights: Dict[str, float] = {}
        self.options = {
            'length_smooth': False,
            'penalty_smooth': True,
            'obstacle_smooth': False,
            'maxratio_smooth': False,
        }
        self.params = {
            'maxratio': None,
            'maxalpha': None,
            'enlargement': None
        }
        if opts:
            self.options.update(opts)
        self.update_weights()

    def update_weights(self):
        for region_name in self.map.region_names():
            if region_name not in self.weights:
                self.weights[region_name] = 1

    def set_weight(self, region_name: str, w: float):
        assert region_name in self.map.regions
        self.weights[region_name] = w

    # ---- parameter vector / flags of the C-ABI ---------------------------------------------------------
    def flags(self) -> int:
        o = self.options
        return ((_lib.UAM_LENGTH_SMOOTH if o['length_smooth'] else 0) |
                (_lib.UAM_PENALTY_SMOOTH if o['penalty_smooth'] else 0) |
                (_lib.UAM_OBSTACLE_SMOOTH if o['obstacle_smooth'] else 0) |
                (_lib.UAM_MAXRATIO_SMOOTH if o['maxratio_smooth'] else 0))

    def parameter_vector(self, need_constraints: bool = False, weights=None) -> np.ndarray:
        """p = [x_start, x_goal, maxratio, maxalpha, enlargement, w...] in region insertion order
        (solver.py:60-68, main.py:145-150)."""
        self.update_weights()
        prm = self.params
        if prm['enlargement'] is None:
            raise TypeError("params['enlargement'] is None (the reference fails on `h(x) - None`)")
        if need_constraints and (prm['maxratio'] is None or prm['maxalpha'] is None):
            raise TypeError("params['maxratio'] / params['maxalpha'] must be set for get_nonlincon")
        mr = np.nan if prm['maxratio'] is None else float(prm['maxratio'])
        ma = np.nan if prm['maxalpha'] is None else float(prm['maxalpha'])
        w = [float(self.weights[n]) for n in self.map.region_names()] if weights is None else list(weights)
        xs = np.asarray(self.map.x_start, dtype=np.float64).ravel()
        xg = np.asarray(self.map.x_goal, dtype=np.float64).ravel()
        return np.concatenate([xs, xg, [mr, ma, float(prm['enlargement'])], w]).astype(np.float64)

    def _batched(self, z):
        if _is_tensor(z):
            return (z, False) if z.dim() == 2 else (z.reshape(1, -1), True)
        z = np.asarray(z, dtype=np.float64)
        return (z, False) if z.ndim == 2 else (z.reshape(1, -1), True)

    # ---- the hot path --------------------------------------------------------------------------------------
    def score(self, z, want_g: bool = False):
        """(cost, collide, g) for one path or a batch: get_cost + any_j Map.collides(z_j) + get_nonlincon."""
        Z, single = self._batched(z)
        p = self.parameter_vector(need_constraints=want_g)
        cost, col, g = self.map.engine().score_analytic(Z, self.N, p, self.flags(), want_g)
        if single:
            return float(cost[0]), bool(col[0]), (g[0] if g is not None else None)
        return cost, col, g

    def get_cost(self, z):
        """(N+1) * length_of(z) + sum_j penalty(z_j) / N   (problem.py:38-44)."""
        return self.score(z, want_g=False)[0]

    def get_cost_gradient(self, z):
        """(cost, d cost / d z_) for one path or a batch; z_ layout as get_cost, the gradient has the same shape (its
        columns 2..2N+1 are the solver's decision variables, solver.py:59).  Analytic stand-in for the derivative
        CasADi generates for OpEn (solver.py:82-101).  With length_smooth = False a pair of coincident consecutive points
        (always the case for (map.x_start, z_0) in the reference's layout) contributes the zero subgradient, never NaN."""
        Z, single = self._batched(z)
        cost, grad = self.map.engine().grad_analytic(Z, self.N, self.parameter_vector(), self.flags())
        return (float(cost[0]), grad[0]) if single else (cost, grad)

    def get_nonlincon(self, z):
        """[ratio-hi, ratio-lo, angle] per interior waypoint, then psi_obstacle(z_j) per obstacle and waypoint
        (problem.py:84-114)."""
        return self.score(z, want_g=True)[2]

    def path_collides(self, z):
        """any_j map.collides(z_j) for one path or a batch."""
        return self.score(z, want_g=False)[1]

    def get_total_penalty_function(self) -> Callable:
        """x -> sum over regions of the weighted region penalty (problem.py:49-56)."""
        def total_penalty(x):
            x = np.asarray(x, dtype=np.float64)
            out = self.map.engine().eval_points(x.reshape(-1, 2), self.parameter_vector(), self.flags(), want=('region',))
            pen = np.zeros(out['region'].shape[0])
            for r in range(out['region'].shape[1]):        # `penalty += weighted_psi(x)` in region order
                pen = pen + out['region'][:, r]
            return float(pen[0]) if x.ndim == 1 else pen
        return total_penalty

    def get_penalty_function(self, region_name=None):
        """x -> w * sum_s psi_s(x)/psi_s(center_s); region_name None = the hard obstacles with w = 1 and
        obstacle_smooth (problem.py:59-82)."""
        if region_name is not None and region_name not in self.map.regions:
            raise KeyError(region_name)

        def penalty(x):
            x = np.asarray(x, dtype=np.float64)
            eng = self.map.engine()
            p = self.parameter_vector()
            if region_name is None:
                v = eng.eval_points(x.reshape(-1, 2), p, self.flags(), want=('obstacle',))['obstacle']
            else:
                r = self.map.region_names().index(region_name)
                v = eng.eval_points(x.reshape(-1, 2), p, self.flags(), want=('region',))['region'][:, r]
            return float(v[0]) if x.ndim == 1 else v
        return penalty

    def length_of(self, x, smooth=False):
        """sum_{k=0}^{N} nrm(y_{k+1} - y_k), y = [map.x_start; x; map.x_goal]  (problem.py:130-146).
        x: flat (2M,) or batch (B, 2M); only the first N+1 pairs count, whatever M is."""
        X, single = self._batched(x)
        out = self.map.engine().length_of(X, self.N, self.map.x_start, self.map.x_goal, bool(smooth))
        return float(out[0]) if single else out

"""ctypes loader of oracle/libuam_oracle_c.so -- the C / OpenMP restatement of ``uam_oracle.score_paths_raster`` and
``uam_oracle.grid_search`` (TEST INFRASTRUCTURE ONLY: full-size parity checks and the all-cores CPU baseline of
bench.py; see oracle/uam_oracle_c.c).  ``build()`` runs oracle/Makefile (gcc)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libuam_oracle_c.so')
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'uam_oracle_c.c')
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        r = subprocess.run(['make', '-C', _HERE, '-B', 'libuam_oracle_c.so'], capture_output=True, text=True)
        if r.returncode != 0:       # no libgomp: single-threaded build
            r = subprocess.run(['make', '-C', _HERE, '-B', 'libuam_oracle_c.so', 'OPENMP='], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('oracle/Makefile failed:\n' + r.stdout + r.stderr)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        vp, i, d, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
        lib.uam_oc_score_paths_raster.restype = i
        lib.uam_oc_score_paths_raster.argtypes = [vp, vp, i, i, i, d, d, d, d, vp, i64, i, vp, d, i, vp, vp, vp, vp, i]
        lib.uam_oc_grid_search.restype = i
        lib.uam_oc_grid_search.argtypes = [vp, vp, i, i, i, vp, i, vp, vp, i]
        lib.uam_oc_num_threads.restype = i
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def num_threads() -> int:
    return int(load().uam_oc_num_threads())


def score_paths_raster(layers, occ, geo, Z, weights, samples_per_cell=0.0, length_smooth=True, x_start=None, threads=0):
    """Same contract as uam_oracle.score_paths_raster -> (cost f64 (B,), collide bool (B,), n_samples int64 (B,))."""
    layers = np.ascontiguousarray(layers, dtype=np.float32)
    occ = None if occ is None else np.ascontiguousarray(occ, dtype=np.uint8)
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    L, H, W = layers.shape
    B, Wp = Z.shape[0], Z.shape[1] // 2
    w = np.ascontiguousarray(weights, dtype=np.float64)
    assert w.shape[0] == L
    xs = None if x_start is None else np.ascontiguousarray(x_start, dtype=np.float64)
    cost = np.empty(B, dtype=np.float64)
    col = np.empty(B, dtype=np.uint8)
    ns = np.empty(B, dtype=np.int64)
    rc = load().uam_oc_score_paths_raster(_p(layers), _p(occ), L, H, W, *[float(g) for g in geo], _p(Z), B, Wp, _p(w),
                                          float(samples_per_cell), int(bool(length_smooth)), _p(xs), _p(cost), _p(col), _p(ns),
                                          int(threads))
    assert rc == 0
    return cost, col.astype(bool), ns


def grid_search(cost, sources, blocked=None, want_parent=True, threads=0):
    """cost (H,W) | (bands,H,W) uint16, sources (Q,2) | (Q,3) -> dist (Q,...) int64, parent (Q,...) int32 | None."""
    cost = np.ascontiguousarray(cost, dtype=np.uint16)
    flat2d = cost.ndim == 2
    shape = cost.shape
    c3 = cost[None] if flat2d else cost
    b3 = None if blocked is None else np.ascontiguousarray(blocked, dtype=np.uint8).reshape(c3.shape)
    src = np.asarray(sources, dtype=np.int32).reshape(-1, 2 if flat2d else 3)
    if flat2d:
        src = np.concatenate([np.zeros((src.shape[0], 1), dtype=np.int32), src], axis=1)
    src = np.ascontiguousarray(src)
    Q = src.shape[0]
    dist = np.empty((Q,) + shape, dtype=np.int64)
    parent = np.empty((Q,) + shape, dtype=np.int32) if want_parent else None
    rc = load().uam_oc_grid_search(_p(c3), _p(b3), c3.shape[0], c3.shape[1], c3.shape[2], _p(src), Q, _p(dist), _p(parent), int(threads))
    assert rc == 0
    return dist, parent

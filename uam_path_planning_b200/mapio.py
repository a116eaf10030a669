"""Safe reader of the planner's on-disk map format.

``map_generation/data_manager.py:56-72`` writes ``vertices = [polygon([x, y], ...), ...]`` as Python source and
``path_generation/utils.py:29-35`` reads it back with ``exec``.  This reader accepts the same text but walks
its syntax tree instead of executing it: only an assignment of a list of ``polygon`` / ``ball`` / ``square``
calls with numeric-literal arguments is allowed.
"""
from __future__ import annotations

import ast
from typing import List

from .shapes import QuadraticObstacle, ball, polygon, square

_CTORS = {'polygon': polygon, 'ball': ball, 'square': square}


def _literal(node):
    if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
        return node.value
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
        v = _literal(node.operand)
        return -v if isinstance(node.op, ast.USub) else v
    if isinstance(node, (ast.List, ast.Tuple)):
        return [_literal(e) for e in node.elts]
    raise ValueError(f'unsupported expression in map file: {ast.dump(node)[:80]}')


def parse_shapes(text: str, varname: str = 'vertices') -> List[QuadraticObstacle]:
    tree = ast.parse(text)
    for stmt in tree.body:
        if not (isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name)):
            raise ValueError('map file may only contain assignments')
        if stmt.targets[0].id != varname:
            continue
        if not isinstance(stmt.value, ast.List):
            raise ValueError(f'{varname} must be a list of shapes')
        shapes = []
        for call in stmt.value.elts:
            if not (isinstance(call, ast.Call) and isinstance(call.func, ast.Name) and call.func.id in _CTORS
                    and not call.keywords):
                raise ValueError('map file entries must be polygon(...), ball(...) or square(...)')
            shapes.append(_CTORS[call.func.id](*[_literal(a) for a in call.args]))
        return shapes
    raise KeyError(varname)


def get_var_from_file(filename: str, varname: str = 'vertices') -> List[QuadraticObstacle]:
    """Same call as path_generation/utils.py:29-35, without exec."""
    with open(filename, 'r') as fh:
        return parse_shapes(fh.read(), varname)


def save_polygons(polygons, output_file: str) -> None:
    """Writer of the same on-disk format (map_generation/data_manager.py:56-72): ``vertices = [polygon([x, y], ...),``
    one polygon per line, coordinates divided by 1000 (metres -> km) and printed with ``str`` like the reference.
    `polygons`: sequences of (x, y) vertices in METRES (a closing vertex equal to the first is dropped)."""
    lines = []
    for poly in polygons:
        coords = [tuple(map(float, c)) for c in poly]
        if len(coords) > 1 and coords[0] == coords[-1]:
            coords = coords[:-1]
        lines.append('polygon(' + ', '.join('[' + str(x / 1000) + ', ' + str(y / 1000) + ']' for x, y in coords) + ')')
    with open(output_file, 'w') as fh:
        fh.write('vertices = [' + ',\n'.join(lines) + ('\n' if lines else '') + ']')


# path_generation/main.py:103-116: start / goal of the shipped scenario in EPSG:2443 metres
DEFAULT_START_POINT = [35590.685, -27711.422]
DEFAULT_END_POINT = [26478.673, 9564.082]


def result_points(x, start_point=None, end_point=None):
    """The point list both exporters of the reference build (make_result_line_shp / save_points_to_shp,
    path_generation/main.py:103-116): [start_point] + [(1000 x_i, 1000 y_i) for the interior waypoints] + [end_point],
    in EPSG:2443 metres.  `x` is the flat solver vector (2N,) in km.  Returns an (N + 2, 2) float64 array."""
    import numpy as np
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if x.size % 2:
        raise ValueError('x must hold interleaved (x, y) pairs')
    sp = DEFAULT_START_POINT if start_point is None else start_point
    ep = DEFAULT_END_POINT if end_point is None else end_point
    pts = [list(map(float, sp))] + [[1000 * x[i], 1000 * x[i + 1]] for i in range(0, len(x), 2)] + [list(map(float, ep))]
    return np.asarray(pts, dtype=np.float64)


def result_wkt(x, start_point=None, end_point=None, kind: str = 'line') -> str:
    """WKT of the best path in EPSG:2443 metres: 'line' -> LINESTRING (make_result_line_shp), 'points' -> MULTIPOINT
    (save_points_to_shp).  The reference then reprojects to EPSG:4612 and writes a shapefile through geopandas /
    pyproj / fiona, none of which exist here: this is the array / text hand-off to that step."""
    pts = result_points(x, start_point, end_point)
    body = ', '.join(f'{p[0]!r} {p[1]!r}' for p in pts.tolist())
    if kind == 'line':
        return f'LINESTRING ({body})'
    if kind == 'points':
        return 'MULTIPOINT (' + ', '.join(f'({p[0]!r} {p[1]!r})' for p in pts.tolist()) + ')'
    raise ValueError("kind must be 'line' or 'points'")

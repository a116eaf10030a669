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

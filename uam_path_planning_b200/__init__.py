"""uam_path_planning_b200: B200-native scorer for candidate flight paths (cost + collision) behind the Python
entry points of nomaporon/uam_path_planning's path_generation package.

numpy arrays in -> ctypes C-ABI (include/uam_b200.h, libuam_b200.so) -> hand-written CUDA for sm_100a.
No CPU fallback: the shape constructors and containers are host data, every evaluation runs on the GPU.
"""
from ._lib import UamError
from .shapes import QuadraticObstacle, Inequality, polygon, ball, square
from .region_map import Map, RegionMap
from .problem import Problem
from .solver import Solver
from .mapio import get_var_from_file, parse_shapes, save_polygons, result_points, result_wkt
from .engine import Engine, default_engine
from .raster import RasterMap, load_dem_mask
from . import distributed
from . import mapgen

__all__ = ['UamError', 'QuadraticObstacle', 'Inequality', 'polygon', 'ball', 'square', 'Map', 'RegionMap', 'Problem',
           'Solver', 'get_var_from_file', 'parse_shapes', 'save_polygons', 'result_points', 'result_wkt', 'Engine', 'default_engine', 'RasterMap', 'load_dem_mask',
           'distributed', 'mapgen']

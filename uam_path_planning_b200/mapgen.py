"""GPU front-end of map_generation: DEM band -> land / sea mask -> connected regions -> minimum-area rectangles.

The reference does this through GDAL / GEOS / OpenCV on the host (map_generation/data_manager.py:11-19,
data_processor.py:16-34,67-71, main.py:21-38):

    mask     = image > threshold                         (or image == -9999)
    polygons = rasterio.features.shapes(mask)            one polygon per 4-connected region, holes as interior rings
    polygons = [p for p in unary_union(polygons).geoms if p.area > min_area]
    rects    = [Polygon(np.intp(cv2.boxPoints(cv2.minAreaRect(exterior ring)))) for p in polygons]      (p.area <= large_area)
    rects    = [r for r in rects if r.area > min_approx_polygon_area]

Here the raster steps run on the GPU (`Engine.dem_mask / label_components / component_stats / component_rects`); what comes
back is one small table of rectangles.  Regions touching only at a corner stay separate polygons in the reference as well
(`unary_union` does not merge polygons that share a single point), so a 4-connected component = a reference polygon, its
cell count * cell area = `polygon.area`, and the hull of its cell corners = the hull of the exterior ring.

Not covered: the reference cuts polygons larger than `large_area` into 5 x 5 boxes before approximating
(data_processor.py:36-53; box edges fall inside cells, GEOS clips the rings); such components are returned whole with
`large = True` so the caller can treat them as it likes.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import Engine, _is_tensor


def rect_area(rect: np.ndarray) -> np.ndarray:
    """Shoelace area of (K,4,2) corner arrays."""
    x, y = rect[..., 0], rect[..., 1]
    return 0.5 * np.abs(np.sum(x * np.roll(y, -1, axis=-1) - np.roll(x, -1, axis=-1) * y, axis=-1))


def dem_rectangles(image, geo, threshold_dem: float = 0.0, min_area: float = 750000.0, large_area: float = 32000000.0,
                   min_approx_polygon_area: float = 780000.0, connectivity: int = 4, engine: Optional[Engine] = None,
                   integer: bool = True):
    """DataManager.load_dem_polygons_from_geotiff + DataProcessor.process_polygons (without the large-polygon split) on the
    GPU.  image: (H,W) float32 band (numpy or CUDA tensor); geo = (x0, dx, y0, dy) of the cell corners' affine
    (rasterio's transform: x = x0 + col dx, y = y0 + row dy; metres).  Areas in the units of geo squared.

    Returns a dict: rects (K,4,2) -- int64 like np.intp(cv2.boxPoints(..)) (truncation toward zero) when `integer`, else
    float64 --, labels of the K components, their areas, `large` flags, and n_components / n_polygons_over_min_area."""
    import torch
    eng = engine or Engine()
    if not _is_tensor(image):
        image = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(f'cuda:{eng.device}')
    mask = eng.dem_mask(image.contiguous(), threshold_dem)
    labels, n = eng.label_components(mask, connectivity)
    x0, dx, y0, dy = [float(v) for v in geo]
    cell_area = abs(dx * dy)
    out = {'rects': np.zeros((0, 4, 2), dtype=np.int64 if integer else np.float64), 'labels': np.zeros(0, dtype=np.int32),
           'area': np.zeros(0), 'large': np.zeros(0, dtype=bool), 'n_components': n, 'n_polygons_over_min_area': 0}
    if n == 0:
        return out
    area, bbox = eng.component_stats(labels, n)
    a = area.to(torch.float64) * cell_area
    ids = (torch.nonzero(a > min_area).reshape(-1) + 1).to(torch.int32)          # p.area > self.min_area
    out['n_polygons_over_min_area'] = int(ids.numel())
    if ids.numel() == 0:
        return out
    rect = eng.component_rects(labels, n, bbox, ids, (x0, dx, y0, dy)).cpu().numpy()
    if integer:
        rect = np.trunc(rect).astype(np.int64)           # box1 = np.intp(box1)
    keep = rect_area(rect.astype(np.float64)) > min_approx_polygon_area          # p.area > self.min_approx_polygon_area
    ids_h = ids.cpu().numpy()
    a_h = a.cpu().numpy()[ids_h - 1]
    out.update(rects=rect[keep], labels=ids_h[keep], area=a_h[keep], large=(a_h > large_area)[keep])
    return out

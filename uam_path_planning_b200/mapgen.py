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

Polygons larger than `large_area` are cut like the reference does (data_processor.py:25-27,34-53): a `divisions` x
`divisions` grid of boxes over the polygon's bounding box, every piece of polygon.intersection(box) approximated by its own
rectangle.  The box edges fall inside cells; on the grid refined `divisions` times they are cell boundaries, so a box is an
exact block of sub-cells (`Engine.component_submask`), its 4-connected regions are the pieces GEOS would return, and the
hull of a piece's sub-cell corners is the hull of the clipped exterior ring -- the same labelling / rectangle kernels run
on each box.  Pieces come in the reference's box order (x index outer, y index inner, data_processor.py:39-41), inside a
box in raster-scan order of their first cell.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .engine import Engine, _is_tensor


def rect_area(rect: np.ndarray) -> np.ndarray:
    """Shoelace area of (K,4,2) corner arrays."""
    x, y = rect[..., 0], rect[..., 1]
    return 0.5 * np.abs(np.sum(x * np.roll(y, -1, axis=-1) - np.roll(x, -1, axis=-1) * y, axis=-1))


def split_component_rects(eng: Engine, labels, label: int, bbox, geo, divisions: int = 5):
    """DataProcessor._divide_and_approximate_polygon (data_processor.py:34-53) for one component: (P,4,2) float64 corner
    arrays of the minimum-area rectangles of the pieces, and the (P,) box index j * divisions + k of each."""
    x0, dx, y0, dy = [float(v) for v in geo]
    r0, r1, c0, c1 = [int(v) for v in bbox]
    nr, nc = r1 - r0 + 1, c1 - c0 + 1
    rects, boxes = [], []
    # the reference's k counts boxes from miny upwards: that is from the last row block when dy < 0
    rows = list(range(divisions)) if dy > 0 else list(range(divisions - 1, -1, -1))
    cols = list(range(divisions)) if dx > 0 else list(range(divisions - 1, -1, -1))
    for j, bc in enumerate(cols):
        for k, br in enumerate(rows):
            sub = eng.component_submask(labels, label, (r0, r1, c0, c1), divisions, br, bc)
            lab, n = eng.label_components(sub, 4)
            if n == 0:
                continue
            _, bb = eng.component_stats(lab, n)
            import torch
            ids = torch.arange(1, n + 1, dtype=torch.int32, device=lab.device)
            sub_geo = (x0 + (c0 + bc * nc / divisions) * dx, dx / divisions, y0 + (r0 + br * nr / divisions) * dy, dy / divisions)
            rects.append(eng.component_rects(lab, n, bb, ids, sub_geo).cpu().numpy())
            boxes.append(np.full(n, j * divisions + k, dtype=np.int32))
    if not rects:
        return np.zeros((0, 4, 2)), np.zeros(0, dtype=np.int32)
    return np.concatenate(rects), np.concatenate(boxes)


def dem_rectangles(image, geo, threshold_dem: float = 0.0, min_area: float = 750000.0, large_area: float = 32000000.0,
                   min_approx_polygon_area: float = 780000.0, connectivity: int = 4, engine: Optional[Engine] = None,
                   integer: bool = True, divisions: int = 5):
    """DataManager.load_dem_polygons_from_geotiff + DataProcessor.process_polygons on the GPU.  image: (H,W) float32 band
    (numpy or CUDA tensor); geo = (x0, dx, y0, dy) of the cell corners' affine (rasterio's transform: x = x0 + col dx,
    y = y0 + row dy; metres).  Areas in the units of geo squared.

    Returns a dict: rects (K,4,2) -- int64 like np.intp(cv2.boxPoints(..)) when `integer` (the corners pass through float32
    first, as cv2 returns them, then truncate toward zero), else float64 --, labels = the component each rectangle comes
    from, area = that component's area, large = the component was split (data_processor.py:25-27), box = index j *
    divisions + k of the piece's box (-1 for an unsplit component), and n_components / n_polygons_over_min_area."""
    import torch
    eng = engine or Engine()
    if not _is_tensor(image):
        image = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(f'cuda:{eng.device}')
    mask = eng.dem_mask(image.contiguous(), threshold_dem)
    labels, n = eng.label_components(mask, connectivity)
    x0, dx, y0, dy = [float(v) for v in geo]
    cell_area = abs(dx * dy)
    out = {'rects': np.zeros((0, 4, 2), dtype=np.int64 if integer else np.float64), 'labels': np.zeros(0, dtype=np.int32),
           'area': np.zeros(0), 'large': np.zeros(0, dtype=bool), 'box': np.zeros(0, dtype=np.int32), 'n_components': n,
           'n_polygons_over_min_area': 0}
    if n == 0:
        return out
    area, bbox = eng.component_stats(labels, n)
    a = area.to(torch.float64) * cell_area
    ids = (torch.nonzero(a > min_area).reshape(-1) + 1).to(torch.int32)          # p.area > self.min_area
    out['n_polygons_over_min_area'] = int(ids.numel())
    if ids.numel() == 0:
        return out
    ids_h = ids.cpu().numpy()
    a_h = a.cpu().numpy()[ids_h - 1]
    bbox_h = bbox.cpu().numpy()
    large = a_h > large_area                                                      # polygon.area > self.large_area
    rect_all, lab_all, area_all, large_all, box_all = [], [], [], [], []
    whole = torch.from_numpy(ids_h[~large]).to(ids.device)
    rect_whole = eng.component_rects(labels, n, bbox, whole, (x0, dx, y0, dy)).cpu().numpy() if whole.numel() else np.zeros((0, 4, 2))
    w = 0
    for i, cid in enumerate(ids_h):                                               # the reference's polygon order = label order
        if large[i]:
            r, b = split_component_rects(eng, labels, int(cid), bbox_h[cid - 1], (x0, dx, y0, dy), divisions)
        else:
            r, b = rect_whole[w:w + 1], np.full(1, -1, dtype=np.int32)
            w += 1
        rect_all.append(r)
        box_all.append(b)
        lab_all.append(np.full(len(r), cid, dtype=np.int32))
        area_all.append(np.full(len(r), a_h[i]))
        large_all.append(np.full(len(r), bool(large[i])))
    rect = np.concatenate(rect_all)
    if integer:
        rect = np.trunc(rect.astype(np.float32)).astype(np.int64)     # cv2.boxPoints returns float32; box1 = np.intp(box1)
    keep = rect_area(rect.astype(np.float64)) > min_approx_polygon_area          # p.area > self.min_approx_polygon_area
    out.update(rects=rect[keep], labels=np.concatenate(lab_all)[keep], area=np.concatenate(area_all)[keep],
               large=np.concatenate(large_all)[keep], box=np.concatenate(box_all)[keep])
    return out

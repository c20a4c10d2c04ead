"""Oracle (test infrastructure): tile grid and crop/write-window arithmetic.

Restates, in plain numpy/Python float64, the reference's
  * ``flair_zonal_detection/slicing.py:51-112``  (generate_patches_from_reference grid loop)
  * ``flair_zonal_detection/inference.py:300,318-343`` (margin crop + output window + clipping)
  * ``flair_zonal_detection/dataset.py:89-117``  (tile read window, boundless zero fill)
on an in-memory georeferenced raster (rasterio / geopandas / shapely are absent in this
image, see oracle/__init__.py).  Pinned by the known answers of SURVEY.md H7
(tests/test_grid.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class Georef:
    """North-up raster georeferencing: what ``rasterio`` exposes as transform/bounds/shape."""
    left: float
    top: float
    res: float
    width: int
    height: int

    @property
    def bounds(self) -> Tuple[float, float, float, float]:
        # rasterio.transform.array_bounds(height, width, from_origin(left, top, res, res)):
        #   w, n = c, f ;  e, s = transform * (width, height)
        e = self.res * self.width + 0.0 * self.height + self.left
        s = 0.0 * self.width + (-self.res) * self.height + self.top
        return (self.left, s, e, self.top)  # (left, bottom, right, top)


def crop_to_geozone(geo: Georef, geozone_bbox: Optional[Sequence[float]]) -> Optional[Georef]:
    """``rasterio.mask.mask(src, shapes, crop=True)`` restricted to what slicing.py uses of it:
    the transform and shape of the cropped array (slicing.py:41-48).

    ``geozone_bbox`` = (minx, miny, maxx, maxy) of the geozone geometries, or None for
    "the whole raster".  Returns None where rasterio raises ValueError (no overlap), which the
    reference turns into an empty GeoDataFrame (slicing.py:43-44).
    Window rule restated from rasterio 1.4.3 ``features.geometry_window`` (pad=0): floor the
    min row/col, ceil the max row/col of the shapes' bounds in pixel space, intersect with
    the raster.
    """
    if geozone_bbox is None:
        return geo
    minx, miny, maxx, maxy = geozone_bbox
    col_a = (minx - geo.left) / geo.res
    col_b = (maxx - geo.left) / geo.res
    row_a = (geo.top - maxy) / geo.res
    row_b = (geo.top - miny) / geo.res
    col_start, col_stop = int(math.floor(min(col_a, col_b))), int(math.ceil(max(col_a, col_b)))
    row_start, row_stop = int(math.floor(min(row_a, row_b))), int(math.ceil(max(row_a, row_b)))
    col_start, row_start = max(col_start, 0), max(row_start, 0)
    col_stop, row_stop = min(col_stop, geo.width), min(row_stop, geo.height)
    if col_stop <= col_start or row_stop <= row_start:
        return None
    return Georef(left=geo.left + col_start * geo.res, top=geo.top - row_start * geo.res,
                  res=geo.res, width=col_stop - col_start, height=row_stop - row_start)


TILE_COLUMNS = ("id", "input_id", "output_id", "job_done", "left", "bottom", "right", "top",
                "left_o", "bottom_o", "right_o", "top_o", "geometry")


def generate_patches(patch_size: int, margin: int, resolution: float, geo: Georef,
                     geozone_bbox: Optional[Sequence[float]] = None,
                     img_path: str = "", output_name: str = "") -> List[Dict]:
    """slicing.py:51-112, verbatim control flow.  ``geometry`` is the full tile's
    ``(minx, miny, maxx, maxy)`` (what ``row.geometry.bounds`` returns in dataset.py:177)."""
    cropped = crop_to_geozone(geo, geozone_bbox)
    if cropped is None:
        return []
    left_overall, bottom_overall, right_overall, top_overall = cropped.bounds
    ref_left_overall, ref_bottom_overall, _, _ = geo.bounds

    geo_output_size = (patch_size * resolution, patch_size * resolution)
    geo_margin = (margin * resolution, margin * resolution)
    geo_step = ((patch_size - 2 * margin) * resolution, (patch_size - 2 * margin) * resolution)

    min_x, min_y = left_overall, bottom_overall
    max_x, max_y = right_overall, top_overall

    tiles: List[Dict] = []
    existing = set()
    for x_coord in np.arange(min_x - geo_margin[0], max_x + geo_margin[0], geo_step[0]):
        for y_coord in np.arange(min_y - geo_margin[1], max_y + geo_margin[1], geo_step[1]):
            if x_coord + geo_output_size[0] > max_x + geo_margin[0]:
                x_coord = max_x + geo_margin[0] - geo_output_size[0]
            if y_coord + geo_output_size[1] > max_y + geo_margin[1]:
                y_coord = max_y + geo_margin[1] - geo_output_size[1]

            left = x_coord + geo_margin[0]
            right = min(x_coord + geo_output_size[0] - geo_margin[0], max_x)
            bottom = y_coord + geo_margin[1]
            top = min(y_coord + geo_output_size[1] - geo_margin[1], max_y)

            key = tuple(round(float(v), 6) for v in (left, bottom, right, top))
            if key in existing:
                continue
            existing.add(key)

            col = int((x_coord - ref_left_overall) // resolution) + 1
            row = int((y_coord - ref_bottom_overall) // resolution) + 1

            if right - left > 0 and top - bottom > 0:
                tiles.append({
                    "id": f"{1}-{row}-{col}",
                    "input_id": img_path,
                    "output_id": output_name,
                    "job_done": 0,
                    "left": float(left), "bottom": float(bottom),
                    "right": float(right), "top": float(top),
                    "left_o": left_overall, "bottom_o": bottom_overall,
                    "right_o": right_overall, "top_o": top_overall,
                    # shapely box(x, y+size, x+size, y).bounds
                    "geometry": (float(x_coord), float(y_coord),
                                 float(x_coord + geo_output_size[0]),
                                 float(y_coord + geo_output_size[1])),
                })
    return tiles


def read_window_px(tile: Dict, geo: Georef, patch_size: int) -> Tuple[int, int]:
    """Pixel (row0, col0) of the tile's full 512x512 read window (dataset.py:97
    ``from_bounds(*bounds, transform)``; the sub-pixel float window offsets (<=1.4e-7 px,
    SURVEY appendix A) are rounded).  May be negative / exceed the raster: the read is
    boundless with fill 0 (dataset.py:108-115)."""
    minx, miny, maxx, maxy = tile["geometry"]
    col0 = int(round((minx - geo.left) / geo.res))
    row0 = int(round((geo.top - maxy) / geo.res))
    return row0, col0


def read_tile(raster: np.ndarray, row0: int, col0: int, patch_size: int) -> np.ndarray:
    """Boundless windowed read with zero fill: raster (C,H,W) -> (C,P,P), same dtype."""
    c, h, w = raster.shape
    out = np.zeros((c, patch_size, patch_size), dtype=raster.dtype)
    r0, r1 = max(row0, 0), min(row0 + patch_size, h)
    c0, c1 = max(col0, 0), min(col0 + patch_size, w)
    if r1 > r0 and c1 > c0:
        out[:, r0 - row0:r1 - row0, c0 - col0:c1 - col0] = raster[:, r0:r1, c0:c1]
    return out


def write_window(tile: Dict, image_bounds: Dict[str, float], out_res: float,
                 pred_h: int, pred_w: int) -> Optional[Tuple[int, int, int, int]]:
    """inference.py:318-343: (top_px, left_px, height_px, width_px) of the window a cropped
    prediction of shape (pred_h, pred_w) is written to, or None when the reference skips
    the tile (``height_px <= 0 or width_px <= 0``)."""
    left_px = int(round((tile["left"] - image_bounds["left"]) / out_res))
    top_px = int(round((image_bounds["top"] - tile["top"]) / out_res))
    height_px, width_px = pred_h, pred_w
    img_height = int(round((image_bounds["top"] - image_bounds["bottom"]) / out_res))
    img_width = int(round((image_bounds["right"] - image_bounds["left"]) / out_res))
    if top_px + height_px > img_height:
        height_px = img_height - top_px
    if left_px + width_px > img_width:
        width_px = img_width - left_px
    if height_px <= 0 or width_px <= 0:
        return None
    return top_px, left_px, height_px, width_px


def tile_plan(tiles: List[Dict], geo: Georef, patch_size: int, margin: int,
              out_res: Optional[float] = None) -> np.ndarray:
    """Integer plan, one row per tile: [row0, col0, top_px, left_px, height_px, width_px]
    (read-window origin, then the write window of the margin-cropped prediction;
    height_px = 0 marks a skipped tile).  int32 (n, 6)."""
    out_res = geo.res if out_res is None else out_res
    left, bottom, right, top = geo.bounds
    ib = {"left": left, "bottom": bottom, "right": right, "top": top}
    s = patch_size - 2 * margin
    if abs(out_res - geo.res) > 1e-6:
        # inference.py:303-325: the window has the size of the ZOOMED prediction; scipy.ndimage.zoom's output shape
        # is round(in * zoom)
        s = int(round(s * (geo.res / out_res)))
    plan = np.zeros((len(tiles), 6), dtype=np.int32)
    for i, t in enumerate(tiles):
        row0, col0 = read_window_px(t, geo, patch_size)
        w = write_window(t, ib, out_res, s, s)
        plan[i, 0:2] = (row0, col0)
        if w is not None:
            plan[i, 2:6] = w
    return plan

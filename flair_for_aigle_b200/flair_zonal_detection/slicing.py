"""Tile grid of the zonal path: drop-in for ``flair_zonal_detection/slicing.py`` of the reference.

``generate_patches_from_reference(config, img_path, geozone)`` keeps the reference's name,
arguments and column layout (slicing.py:20-121) and must return the *same tiles in the same
order* (bit-exact float64 bounds).  The grid is separable, so the axes are generated
independently (same float64 expressions as slicing.py:69-81, same dedupe rule as :83-87) and
combined x-outer / y-inner; ``tests/test_grid.py`` checks equality with the loop restatement
in ``oracle/grid.py`` over the known answers and random shapes.

Also provides the integer plan the CUDA kernels consume (read-window origins, write windows,
ownership windows implementing "last writer wins" of inference.py:343-352).
"""
from __future__ import annotations

import logging
import math
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from .raster import ZoneRaster, open_raster

logger = logging.getLogger(__name__)

try:  # pandas is the table type the reference's callers index (tiles_gdf.iloc[...])
    import pandas as pd
except ImportError:  # pragma: no cover
    pd = None


class TileBox:
    """Stand-in for the shapely box stored in the reference's ``geometry`` column: only
    ``.bounds`` (minx, miny, maxx, maxy) is used downstream (dataset.py:177)."""
    __slots__ = ("bounds",)

    def __init__(self, minx, miny, maxx, maxy):
        self.bounds = (minx, miny, maxx, maxy)

    def __repr__(self):
        return "TileBox(%r, %r, %r, %r)" % self.bounds


def create_box_from_bounds(x_min: float, x_max: float, y_min: float, y_max: float) -> TileBox:
    """slicing.py:13-17 (shapely ``box(x_min, y_max, x_max, y_min)``; bounds are order-free)."""
    return TileBox(min(x_min, x_max), min(y_min, y_max), max(x_min, x_max), max(y_min, y_max))


def _geozone_bbox(geozone) -> Optional[Sequence[float]]:
    """Bounding box of the geozone geometries.  Accepts None (whole raster), a 4-tuple
    (minx, miny, maxx, maxy), anything with ``.bounds`` / ``.total_bounds`` (shapely /
    geopandas), or an iterable of those."""
    if geozone is None:
        return None
    if hasattr(geozone, "total_bounds"):
        return tuple(float(v) for v in geozone.total_bounds)
    if hasattr(geozone, "bounds") and not isinstance(geozone, (list, tuple)):
        b = geozone.bounds
        return tuple(float(v) for v in b)
    if isinstance(geozone, (list, tuple, np.ndarray)):
        if len(geozone) == 4 and all(isinstance(v, (int, float, np.floating, np.integer)) for v in geozone):
            return tuple(float(v) for v in geozone)
        boxes = [_geozone_bbox(g) for g in geozone]
        boxes = [b for b in boxes if b is not None]
        if not boxes:
            return None
        return (min(b[0] for b in boxes), min(b[1] for b in boxes), max(b[2] for b in boxes),
                max(b[3] for b in boxes))
    raise TypeError(f"unsupported geozone description: {type(geozone)!r}")


def _crop_bounds(r: ZoneRaster, bbox: Optional[Sequence[float]]):
    """Bounds of ``rasterio.mask.mask(src, shapes, crop=True)``'s output (slicing.py:41-48):
    the shapes' bounding window, floor/ceil'ed to pixels and clipped to the raster.
    Returns None when they do not overlap (the reference returns an empty frame)."""
    left, bottom, right, top = r.bounds
    if bbox is None:
        return left, bottom, right, top
    res = r.res_value
    ca, cb = (bbox[0] - r.left) / res, (bbox[2] - r.left) / res
    ra, rb = (r.top - bbox[3]) / res, (r.top - bbox[1]) / res
    c0, c1 = max(int(math.floor(min(ca, cb))), 0), min(int(math.ceil(max(ca, cb))), r.width)
    r0, r1 = max(int(math.floor(min(ra, rb))), 0), min(int(math.ceil(max(ra, rb))), r.height)
    if c1 <= c0 or r1 <= r0:
        return None
    w, h = c1 - c0, r1 - r0
    new_left, new_top = r.left + c0 * res, r.top - r0 * res
    return new_left, 0.0 * w + (-res) * h + new_top, res * w + 0.0 * h + new_left, new_top


def _axis(lo: float, hi: float, size: float, gm: float, step: float):
    """One axis of the grid: tile origins (clamped), inner [a, b] intervals, first-occurrence
    dedupe on the 6-decimal rounded interval, then the ``b - a > 0`` filter."""
    origins = np.arange(lo - gm, hi + gm, step)
    out = []
    seen = set()
    for o in origins:
        if o + size > hi + gm:
            o = hi + gm - size
        a = o + gm
        b = min(o + size - gm, hi)
        key = (round(float(a), 6), round(float(b), 6))
        if key in seen:
            continue
        seen.add(key)
        out.append((float(o), float(a), float(b)))
    return out


def generate_patches_from_reference(config: Dict, img_path, geozone_contour_geometries=None):
    """Slice the reference raster into overlapping tiles (slicing.py:20-121).

    Returns a pandas DataFrame with the reference's columns
    ``id,input_id,output_id,job_done,left,bottom,right,top,left_o,bottom_o,right_o,top_o,geometry``
    (empty frame if the geozone misses the raster).  ``geometry`` holds a ``TileBox``."""
    patch_size = config["img_pixels_detection"]
    margin = config["margin"]
    output_name = config["output_name"]
    resolution = config["reference_resolution"]

    src = open_raster(img_path)
    crop = _crop_bounds(src, _geozone_bbox(geozone_contour_geometries))
    if crop is None:
        return pd.DataFrame() if pd is not None else []
    left_o, bottom_o, right_o, top_o = crop
    ref_left, ref_bottom, _, _ = src.bounds

    size = patch_size * resolution
    gm = margin * resolution
    step = (patch_size - 2 * margin) * resolution

    xs = _axis(left_o, right_o, size, gm, step)
    ys = _axis(bottom_o, top_o, size, gm, step)
    # a pair is dropped only when BOTH axis intervals are empty-or-negative-free; the reference
    # filters on (right-left > 0 and top-bottom > 0) after the dedupe
    rows: List[Dict] = []
    img_name = img_path if isinstance(img_path, str) else getattr(img_path, "name", "<memory>")
    for (x, l, r) in xs:
        col = int((x - ref_left) // resolution) + 1
        for (y, b, t) in ys:
            if not (r - l > 0 and t - b > 0):
                continue
            row = int((y - ref_bottom) // resolution) + 1
            rows.append({
                "id": f"{1}-{row}-{col}", "input_id": img_name, "output_id": output_name, "job_done": 0,
                "left": l, "bottom": b, "right": r, "top": t,
                "left_o": left_o, "bottom_o": bottom_o, "right_o": right_o, "top_o": top_o,
                "geometry": create_box_from_bounds(x, x + size, y, y + size),
            })
    if pd is None:  # pragma: no cover
        return rows
    gdf = pd.DataFrame(rows)
    if config.get("write_dataframe", False) and len(gdf):
        # slicing.py:116-119: the tile boxes and their columns as <output_name>_slicing_job.gpkg
        from .gpkg import write_gpkg
        out = os.path.join(config["output_path"], output_name + "_slicing_job.gpkg")

        def box_ring(g):                                  # shapely.geometry.box's vertex order (counter-clockwise)
            x0, y0, x1, y1 = g.bounds
            return [np.asarray([(x1, y0), (x1, y1), (x0, y1), (x0, y0), (x1, y0)], dtype=np.float64)]
        write_gpkg(out, (box_ring(g) for g in gdf.geometry),
                   {c: gdf[c].to_numpy() for c in gdf.columns if c != "geometry"}, getattr(src, "crs", None))
        logger.info(f"[✓] Saved Sliced Boxes: {out}")
    return gdf


# ------------------------------------------------------------------------------------------
# integer plans for the device kernels
# ------------------------------------------------------------------------------------------
def zoom_map(size: int, scale: float) -> np.ndarray:
    """Source index of every output pixel of ``scipy.ndimage.zoom(prediction, scale, order=0)`` along one axis
    (inference.py:212-226 ``resample_prediction``); -1 where scipy writes its constant fill value 0 instead of a source
    pixel (with mode='constant' the last output coordinate can land a rounding error beyond the last input pixel, e.g. at
    scale 0.5 -- the reference's rasters carry that zero row/column per tile, and so do ours).  scipy itself computes the
    map when present (it is a dependency of the reference), so it is the reference's by construction; otherwise the rule
    is restated without that quirk: output length round(size*scale), corner-aligned coordinates o*(size-1)/(out-1),
    nearest = floor(c + 0.5)."""
    try:
        from scipy.ndimage import zoom
        return np.asarray(zoom(np.arange(1, size + 1, dtype=np.int64), scale, order=0), dtype=np.int32) - 1
    except ImportError:  # pragma: no cover - scipy ships with this image
        out = int(round(size * scale))
        if out <= 1:
            return np.zeros(max(out, 0), np.int32)
        c = np.arange(out, dtype=np.float64) * ((size - 1) / (out - 1))
        return np.clip(np.floor(c + 0.5), 0, size - 1).astype(np.int32)


def tile_plan(tiles_gdf, image_bounds: Dict[str, float], ref_res: float, patch_size: int, margin: int,
              out_res: Optional[float] = None) -> np.ndarray:
    """int32 (n,6): [row0, col0, top_px, left_px, height_px, width_px].

    (row0, col0): pixel origin of the tile's full read window (dataset.py:97 ``from_bounds``;
    boundless, zero fill).  (top_px, left_px, height_px, width_px): where the margin-cropped
    prediction lands, computed exactly like inference.py:318-343 (Python ``round``, clipping at
    the bottom/right raster edge; height_px = 0 marks a tile the reference skips)."""
    out_res = ref_res if out_res is None else out_res
    s = patch_size - 2 * margin
    if abs(out_res - ref_res) > 1e-6:
        s = len(zoom_map(s, ref_res / out_res))          # size of the zoomed prediction (inference.py:303-312)
    n = len(tiles_gdf)
    plan = np.zeros((n, 6), dtype=np.int32)
    if n == 0:
        return plan
    img_h = int(round((image_bounds["top"] - image_bounds["bottom"]) / out_res))
    img_w = int(round((image_bounds["right"] - image_bounds["left"]) / out_res))
    lefts = tiles_gdf["left"].to_numpy()
    tops = tiles_gdf["top"].to_numpy()
    geoms = tiles_gdf["geometry"].to_numpy()
    for i in range(n):
        minx, _, _, maxy = geoms[i].bounds
        plan[i, 0] = int(round((image_bounds["top"] - maxy) / ref_res))
        plan[i, 1] = int(round((minx - image_bounds["left"]) / ref_res))
        left_px = int(round((float(lefts[i]) - image_bounds["left"]) / out_res))
        top_px = int(round((image_bounds["top"] - float(tops[i])) / out_res))
        h = w = s
        if top_px + h > img_h:
            h = img_h - top_px
        if left_px + w > img_w:
            w = img_w - left_px
        if h <= 0 or w <= 0:
            h = w = 0
        elif top_px < 0 or left_px < 0:
            # the reference hands rasterio a window with a negative offset here, which fails
            raise ValueError("raster is smaller than the tile's inner window "
                             f"({patch_size}-2*{margin} px): negative write offset ({top_px}, {left_px})")
        plan[i, 2:6] = (top_px, left_px, h, w)
    return plan


def ownership_windows(plan: np.ndarray) -> np.ndarray:
    """int32 (n,4) [r0, r1, c0, c1]: the part of each tile's write window that is still that
    tile's after every LATER tile has written ("last writer wins", inference.py:343-352).

    Tiles form a product grid (x-outer, y-inner), so ownership is separable: along each axis a
    window loses whatever later windows on that axis cover.  Falls back to an exact 2-D sweep
    when the plan is not a product grid."""
    n = plan.shape[0]
    own = np.zeros((n, 4), dtype=np.int32)
    if n == 0:
        return own
    top, left, h, w = (plan[:, 2].astype(np.int64), plan[:, 3].astype(np.int64), plan[:, 4].astype(np.int64),
                       plan[:, 5].astype(np.int64))
    cols, col_first = np.unique(np.stack([left, w], 1), axis=0, return_index=True)
    rows, row_first = np.unique(np.stack([top, h], 1), axis=0, return_index=True)
    col_order = np.argsort(col_first)  # order of first appearance = enumeration order
    row_order = np.argsort(row_first)
    cols, rows = cols[col_order], rows[row_order]
    product = (len(cols) * len(rows) == n)
    if product:
        k = 0
        for ci in range(len(cols)):
            for ri in range(len(rows)):
                if (left[k], w[k]) != tuple(cols[ci]) or (top[k], h[k]) != tuple(rows[ri]):
                    product = False
                    break
                k += 1
            if not product:
                break
    if product:
        def trim(intervals):
            # intervals in enumeration order; each loses what later ones cover.  Later
            # intervals only ever cover a prefix or a suffix here; keep the largest run left.
            res = []
            for i, (a, ln) in enumerate(intervals):
                lo, hi = int(a), int(a + ln)
                alive = np.ones(max(hi - lo, 0), dtype=bool)
                for (a2, ln2) in intervals[i + 1:]:
                    s0, s1 = max(int(a2), lo), min(int(a2 + ln2), hi)
                    if s1 > s0:
                        alive[s0 - lo:s1 - lo] = False
                idx = np.flatnonzero(alive)
                if idx.size == 0:
                    res.append((lo, lo))
                else:
                    if idx[-1] - idx[0] + 1 != idx.size:
                        return None  # not a single run: use the 2-D sweep
                    res.append((lo + int(idx[0]), lo + int(idx[-1]) + 1))
            return res
        ct, rt = trim([tuple(c) for c in cols]), trim([tuple(r) for r in rows])
        if ct is not None and rt is not None:
            k = 0
            for ci in range(len(cols)):
                for ri in range(len(rows)):
                    own[k] = (rt[ri][0], rt[ri][1], ct[ci][0], ct[ci][1])
                    k += 1
            return own
    raise NotImplementedError("tile plan is not a product grid; ownership needs a sequential write")

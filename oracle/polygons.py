"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the polygonisation stage's integer parts.

Reference: flair_zonal_detection/inference.py:356-407 -- per class ``mask = data == cls`` and
``rasterio.features.shapes(mask.astype(uint8), mask=mask, transform)``; GDAL's polygonize default is 4-connectivity,
one polygon (exterior + holes) per connected component, ``poly.area < min_area`` dropped (:368).  rasterio / shapely are
not in this image, so the component structure is restated with scipy.ndimage.label (cross-shaped structuring element =
4-connectivity) and polygons are checked by rasterising them back (even-odd rule at pixel centres), which does not
depend on how a tracer orders or starts its rings.  Parity unpinned by the reference (it has no tests)."""
import numpy as np
from scipy import ndimage

FOUR = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)


def label_components(raster: np.ndarray) -> np.ndarray:
    """int32 [H,W]: label of a pixel = smallest linear index of its 4-connected same-class component."""
    H, W = raster.shape
    out = np.empty((H, W), np.int64)
    lin = np.arange(H * W, dtype=np.int64).reshape(H, W)
    for cls in np.unique(raster):                                   # inference.py:394, :359
        mask = raster == cls
        lab, n = ndimage.label(mask, structure=FOUR)
        if n == 0:
            continue
        first = ndimage.minimum(lin, labels=lab, index=np.arange(1, n + 1)).astype(np.int64)
        out[mask] = first[lab[mask] - 1]
    return out.astype(np.int32)


def component_table(raster: np.ndarray, labels: np.ndarray):
    """(roots, areas, classes) sorted by root."""
    roots, areas = np.unique(labels, return_counts=True)
    classes = raster.reshape(-1)[roots]
    return roots.astype(np.int32), areas.astype(np.int32), classes.astype(np.uint8)


def rasterize_even_odd(rings, H: int, W: int) -> np.ndarray:
    """bool [H,W]: pixel centres inside the polygon whose rings (exterior + holes, pixel-corner coordinates, closed) are
    given, by the even-odd rule.  Rings are rectilinear or simplified; crossings are counted per pixel row at y + 0.5."""
    inside = np.zeros((H, W), bool)
    xs = np.arange(W) + 0.5
    for y in range(H):
        yc = y + 0.5
        cross = []
        for ring in rings:
            r = np.asarray(ring, np.float64)
            x0, y0, x1, y1 = r[:-1, 0], r[:-1, 1], r[1:, 0], r[1:, 1]
            hit = (y0 <= yc) != (y1 <= yc)
            if hit.any():
                t = (yc - y0[hit]) / (y1[hit] - y0[hit])
                cross.append(x0[hit] + t * (x1[hit] - x0[hit]))
        if cross:
            c = np.sort(np.concatenate(cross))
            inside[y] = (np.searchsorted(c, xs, side="right") % 2) == 1
    return inside


def ring_area2(ring) -> float:
    """Twice the signed shoelace area in pixel-corner coordinates (y down): > 0 for exterior rings walked with the
    component on the right."""
    r = np.asarray(ring, np.float64)
    return float(np.sum(r[:-1, 0] * r[1:, 1] - r[1:, 0] * r[:-1, 1]))

"""Vector polygons from the class raster: the drop-in for ``raster_to_polygons`` (inference.py:375-407 with its worker
``_extract_polygons_for_class``, :356-373), called right after ``inference_and_write`` by the product script
(scripts/run_fast_aigle_segmentation.py:119).

Reference: per class, ``rasterio.features.shapes(mask)`` (GDAL polygonize, 4-connectivity) -> shapely polygon -> drop if
``area < min_area`` -> ``simplify(simplification, preserve_topology=True)`` -> GeoDataFrame(class_id, geometry).

Here: the O(H*W) part, connected-component labelling + per-component areas, runs on the GPU on the raster the head kernel
just wrote (``fz_ccl_label`` / ``fz_ccl_areas`` / ``fz_ccl_table``); only the components that survive the class and area
filters are traced, on the host, by ``fz_trace_rings`` (boundary following + Douglas-Peucker).  geopandas / shapely are
not in this image, so the result is a ``PolygonTable`` of GeoJSON-like geometries (what ``shapes`` yields) with the same
two columns.  Differences from GEOS worth knowing: plain Douglas-Peucker per ring (with the reference's tolerance of half
a pixel it only removes collinear and sub-half-pixel vertices, where the two agree); rings start at their top-left
corner."""
import json
from typing import Dict, List, Optional

import numpy as np
import torch

from ..native import NativeError
from .. import native as nv


class _Geometries:
    """Sequence of GeoJSON-like polygons over flat ring arrays; the Python lists are only built for the rows touched."""

    def __init__(self, xy, ring_off, ring_perm, poly_ring_off):
        self.xy, self.ring_off, self.ring_perm, self.poly_ring_off = xy, ring_off, ring_perm, poly_ring_off

    def __len__(self) -> int:
        return len(self.poly_ring_off) - 1

    def rings(self, i: int):
        """numpy views [n,2] of polygon i's rings: exterior first, then the holes."""
        ks = self.ring_perm[self.poly_ring_off[i]:self.poly_ring_off[i + 1]]
        return [self.xy[self.ring_off[k]:self.ring_off[k + 1]] for k in ks]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return {"type": "Polygon", "coordinates": [r.tolist() for r in self.rings(i)]}

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


class PolygonTable:
    """Rows of {class_id, geometry}: ``geometry[i]`` is a GeoJSON-like dict (type Polygon, coordinates = [exterior,
    holes...], rings closed, CRS coordinates).  ``area`` holds each polygon's area in CRS units (pixel count * res^2)."""

    def __init__(self, class_id: np.ndarray, area: np.ndarray, geometry, crs=None, confidence: Optional[np.ndarray] = None):
        self.class_id, self.area, self.geometry, self.crs = class_id, area, geometry, crs
        self.confidence = confidence     # optional third column (vectorize_segmentation: mean confidence of the row's class)

    def __len__(self) -> int:
        return len(self.geometry)

    def __iter__(self):
        for c, g in zip(self.class_id, self.geometry):
            yield {"class_id": int(c), "geometry": g}

    def to_geojson(self, path: Optional[str] = None) -> dict:
        fc = {"type": "FeatureCollection",
              "features": [{"type": "Feature", "properties": {"class_id": int(c), "area": float(a)}, "geometry": g}
                           for c, a, g in zip(self.class_id, self.area, self.geometry)]}
        if self.crs:
            fc["crs"] = {"type": "name", "properties": {"name": str(self.crs)}}
        if path is not None:
            with open(path, "w") as f:
                json.dump(fc, f)
        return fc

    def to_file(self, path: str, driver: str = "GPKG", layer: Optional[str] = None) -> str:
        """``gdf_results.to_file(raster_results_filepath, driver="GPKG")`` (scripts/run_fast_aigle_segmentation.py:123): the
        frame's two columns -- ``class_id`` and the polygon geometry -- as a GeoPackage layer in the table's CRS
        (``gpkg.write_gpkg``); ``driver="GeoJSON"`` writes ``to_geojson``."""
        if driver == "GeoJSON":
            self.to_geojson(path)
            return path
        if driver != "GPKG":
            raise ValueError(f"driver '{driver}': GPKG and GeoJSON are written")
        from .gpkg import write_gpkg
        geoms = self.geometry
        rings = (geoms.rings(i) for i in range(len(geoms))) if isinstance(geoms, _Geometries) else \
            ([np.asarray(r, dtype=np.float64) for r in g["coordinates"]] for g in geoms)
        columns = {"class_id": np.asarray(self.class_id, dtype=np.int64)}
        if self.confidence is not None:
            columns["confidence"] = np.asarray(self.confidence, dtype=np.float64)
        return write_gpkg(path, rings, columns, self.crs, layer)

    @classmethod
    def read_file(cls, path: str, layer: Optional[str] = None) -> "PolygonTable":
        """``gpd.read_file(path)`` of a file written by ``to_file`` (scripts/run_fast_aigle_segmentation.py:131): class ids,
        geometries and CRS; ``area`` is recomputed from the rings (shoelace, holes subtracted)."""
        from .gpkg import read_gpkg
        cols, geoms, crs = read_gpkg(path, layer)

        def ring_area(r):
            return 0.5 * abs(float(np.dot(r[:-1, 0], r[1:, 1]) - np.dot(r[1:, 0], r[:-1, 1])))
        area = np.asarray([ring_area(g[0]) - sum(ring_area(h) for h in g[1:]) if g else 0.0 for g in geoms])
        geometry = [{"type": "Polygon", "coordinates": [r.tolist() for r in g]} for g in geoms]
        conf = np.asarray(cols["confidence"], dtype=np.float64) if "confidence" in cols else None
        return cls(np.asarray(cols.get("class_id", np.zeros(len(geoms))), dtype=np.int64), area, geometry, crs, conf)

    @classmethod
    def concat(cls, tables: List["PolygonTable"]) -> "PolygonTable":
        """``pd.concat(gdf_results_list, ignore_index=True)`` (scripts/run_fast_aigle_segmentation.py:132)."""
        tables = [t for t in tables if len(t)]
        if not tables:
            return cls(np.zeros(0, np.int64), np.zeros(0), [], None)
        geometry = [g for t in tables for g in t.geometry]
        conf = None
        if all(t.confidence is not None for t in tables):
            conf = np.concatenate([np.asarray(t.confidence, dtype=np.float64) for t in tables])
        return cls(np.concatenate([np.asarray(t.class_id, dtype=np.int64) for t in tables]),
                   np.concatenate([np.asarray(t.area, dtype=np.float64) for t in tables]), geometry, tables[0].crs, conf)


def _device_raster(src, device):
    """-> (uint8 [H,W] CUDA tensor, left, top, res, crs) from a RasterSink, the {task: sink} dict the reference passes,
    a GeoTIFF path, or a (array, left, top, res[, crs]) tuple."""
    from .raster import RasterSink
    if isinstance(src, dict):
        src = src['AERIAL_LABEL-COSIA'] if 'AERIAL_LABEL-COSIA' in src else next(iter(src.values()))   # inference.py:389
    if isinstance(src, RasterSink):
        if src.device_array is not None:
            arr = src.device_array[0]
        else:
            arr = torch.from_numpy(np.ascontiguousarray(src.to_host()[0])).to(device)
        return arr.contiguous(), src.left, src.top, src.res_value, src.crs
    if isinstance(src, (tuple, list)):
        arr, left, top, res = src[:4]
        crs = src[4] if len(src) > 4 else None
        t = arr if torch.is_tensor(arr) else torch.from_numpy(np.ascontiguousarray(arr))
        if t.dim() == 3:
            t = t[0]
        return t.to(device=device, dtype=torch.uint8).contiguous(), float(left), float(top), float(res), crs
    path = getattr(src, 'name', src)
    from .geotiff import read_geotiff
    arr, left, top, res, crs = read_geotiff(str(path))
    return torch.from_numpy(np.ascontiguousarray(arr[0] if arr.ndim == 3 else arr)).to(device), left, top, res, crs


def raster_to_polygons(tiff_path, ignore_background: bool = True, background_value: int = 18, min_area: float = 1.0,
                       simplification: float = 0.1, n_jobs: Optional[int] = None, device=None) -> PolygonTable:
    """inference.py:375-407.  ``n_jobs`` (the reference's per-class process pool) is accepted and ignored: the labelling
    is one GPU pass over all classes.  Rows are ordered by class, then by the component's first pixel (row-major)."""
    if not torch.cuda.is_available():
        raise NativeError("raster_to_polygons runs on CUDA only (no CPU fallback)")
    device = torch.device(device if device is not None else "cuda")
    raster, left, top, res, crs = _device_raster(tiff_path, device)
    H, W = raster.shape
    labels = nv.ccl_label(raster)
    # inference.py:368 (poly.area < min_area -> skip; a component's polygon area is its pixel count * res^2) and :395-396
    min_px = int(np.ceil(min_area / (res * res) - 1e-9))
    roots, areas, classes = nv.ccl_components(raster, labels, min_area_px=min_px,
                                              ignore_class=background_value if ignore_background else -1)
    if roots.size == 0:
        return PolygonTable(np.zeros(0, np.int64), np.zeros(0), [], crs)
    labels_host = labels.cpu().numpy()
    ring_root, ring_hole, ring_off, xy = nv.trace_rings(labels_host, roots, simplification / res if simplification > 0 else 0.0)
    xy[:, 0] = left + xy[:, 0] * res                                     # pixel corners -> CRS (north-up transform)
    xy[:, 1] = top - xy[:, 1] * res
    order = np.lexsort((roots, classes))                                 # rows: by class, then by first pixel
    inv = np.empty_like(order)
    inv[order] = np.arange(order.size)
    ring_poly = inv[np.searchsorted(roots, ring_root)]
    ring_perm = np.lexsort((np.arange(ring_root.size), ring_hole, ring_poly))     # exterior first inside a polygon
    poly_ring_off = np.concatenate([[0], np.cumsum(np.bincount(ring_poly, minlength=order.size))])
    geometry = _Geometries(xy, ring_off, ring_perm, poly_ring_off)
    return PolygonTable(classes[order].astype(np.int64), areas[order].astype(np.float64) * (res * res), geometry, crs)


def class_mean_confidence(labels: np.ndarray, confidence: np.ndarray, class_ids) -> np.ndarray:
    """inference.py:588,610: ``confidence[labels == value].mean()`` -- the mean over ALL pixels of the class, which is what
    every polygon of that class carries in the reference (not a per-polygon mean).  One bincount pass for all classes."""
    lab = np.asarray(labels).reshape(-1).astype(np.int64)
    n = int(lab.max()) + 1 if lab.size else 1
    sums = np.bincount(lab, weights=np.asarray(confidence, dtype=np.float64).reshape(-1), minlength=n)
    counts = np.bincount(lab, minlength=n)
    ids = np.asarray(class_ids, dtype=np.int64)
    return sums[ids] / np.maximum(counts[ids], 1)


def vectorize_segmentation_parallel(labels, confidence, transform, n_jobs: int = 4, device=None, **kwargs) -> PolygonTable:
    """inference.py:598-632 (and ``vectorize_segmentation``, :574-595): polygons of every class but 0 of the label map the
    accumulating ``inference()`` returns, ``min_area`` (default 4.0 CRS units^2) and ``simplification_tolerance`` (default 1.0)
    like the reference, each row carrying its class's mean confidence.  ``transform``: the raster's affine as
    (a, b, c, d, e, f) or anything with those attributes (north-up: b = d = 0).  ``n_jobs`` is accepted and ignored (one GPU
    labelling pass instead of a process per class)."""
    t = tuple(transform)[:6] if not hasattr(transform, "a") else (transform.a, transform.b, transform.c, transform.d,
                                                                   transform.e, transform.f)
    a, b, c, d, e, f = (float(v) for v in t)
    if b != 0.0 or d != 0.0 or abs(a + e) > 1e-9 * abs(a):
        raise NotImplementedError("vectorize_segmentation: north-up rasters with square pixels only")
    labels = np.asarray(labels)
    table = raster_to_polygons((labels.astype(np.uint8), c, f, a, kwargs.get("crs", "EPSG:5490")), ignore_background=True,
                               background_value=0, min_area=kwargs.get("min_area", 4.0),
                               simplification=kwargs.get("simplification_tolerance", 1.0), n_jobs=n_jobs, device=device)
    table.confidence = class_mean_confidence(labels, confidence, table.class_id) if len(table) else np.zeros(0)
    return table


def vectorize_segmentation(labels, confidence, transform, crs="EPSG:5490", simplification_tolerance=1.0, device=None) -> PolygonTable:
    """inference.py:574-595: the same without an area filter."""
    return vectorize_segmentation_parallel(labels, confidence, transform, crs=crs, min_area=0.0,
                                           simplification_tolerance=simplification_tolerance, device=device)

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

    def __init__(self, class_id: np.ndarray, area: np.ndarray, geometry, crs=None):
        self.class_id, self.area, self.geometry, self.crs = class_id, area, geometry, crs

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
        return write_gpkg(path, rings, {"class_id": np.asarray(self.class_id, dtype=np.int64)}, self.crs, layer)

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
        return cls(np.asarray(cols.get("class_id", np.zeros(len(geoms))), dtype=np.int64), area, geometry, crs)

    @classmethod
    def concat(cls, tables: List["PolygonTable"]) -> "PolygonTable":
        """``pd.concat(gdf_results_list, ignore_index=True)`` (scripts/run_fast_aigle_segmentation.py:132)."""
        tables = [t for t in tables if len(t)]
        if not tables:
            return cls(np.zeros(0, np.int64), np.zeros(0), [], None)
        geometry = [g for t in tables for g in t.geometry]
        return cls(np.concatenate([np.asarray(t.class_id, dtype=np.int64) for t in tables]),
                   np.concatenate([np.asarray(t.area, dtype=np.float64) for t in tables]), geometry, tables[0].crs)


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

"""In-memory georeferenced raster: the part of ``rasterio.DatasetReader`` the hot path uses.

The reference opens rasters with rasterio (flair_zonal_detection/dataset.py:44-46,
inference.py:92-101, model_utils.py:11-16); this image has no rasterio/GDAL, and raster file
I/O is the "next" row of SURVEY.md section 8(f).  ``ZoneRaster`` carries what the path needs
-- pixels (C,H,W), bounds, resolution -- and can be built from a numpy array, a ``.npy`` file
(+ ``.json`` sidecar) or, when rasterio is installed, any GDAL-readable file.
"""
from __future__ import annotations

import json
import os
from collections import namedtuple
from typing import Optional, Tuple

import numpy as np

BoundingBox = namedtuple("BoundingBox", ["left", "bottom", "right", "top"])

_REGISTRY = {}


class ZoneRaster:
    def __init__(self, array: np.ndarray, left: float, top: float, res: float, crs: Optional[str] = None,
                 name: str = "<memory>"):
        if array.ndim == 2:
            array = array[None]
        self.array = array
        self.left, self.top, self.res_value, self.crs, self.name = float(left), float(top), float(res), crs, name

    # -- rasterio.DatasetReader look-alikes ------------------------------------------------
    @property
    def count(self) -> int:
        return self.array.shape[0]

    @property
    def height(self) -> int:
        return self.array.shape[1]

    @property
    def width(self) -> int:
        return self.array.shape[2]

    @property
    def shape(self) -> Tuple[int, int]:
        return (self.height, self.width)

    @property
    def res(self) -> Tuple[float, float]:
        return (self.res_value, self.res_value)

    @property
    def bounds(self) -> BoundingBox:
        # rasterio: array_bounds(height, width, transform)
        right = self.res_value * self.width + 0.0 * self.height + self.left
        bottom = 0.0 * self.width + (-self.res_value) * self.height + self.top
        return BoundingBox(self.left, bottom, right, self.top)

    @property
    def profile(self) -> dict:
        return {"driver": "MEM", "dtype": str(self.array.dtype), "count": self.count, "height": self.height,
                "width": self.width, "crs": self.crs, "transform": (self.res_value, 0.0, self.left, 0.0,
                                                                     -self.res_value, self.top)}

    def read(self, indexes=None) -> np.ndarray:
        if indexes is None:
            return self.array
        if isinstance(indexes, int):
            return self.array[indexes - 1]
        return self.array[[i - 1 for i in indexes]]

    def close(self) -> None:
        pass


def register_raster(path: str, raster: ZoneRaster) -> None:
    """Make an in-memory raster addressable by a path string, so configs keep using
    ``input_img_path`` exactly like the reference (tests, synthetic benchmarks)."""
    _REGISTRY[path] = raster


def open_raster(path) -> ZoneRaster:
    """``rasterio.open(path)`` stand-in.  Accepts a ZoneRaster, a registered name, ``*.npy``
    (with ``<path>.json`` = {left, top, res[, crs]}) or, if rasterio is importable, any raster."""
    if isinstance(path, ZoneRaster):
        return path
    if path in _REGISTRY:
        return _REGISTRY[path]
    if isinstance(path, str) and path.endswith(".npy"):
        arr = np.load(path, mmap_mode="r")
        with open(path + ".json") as f:
            meta = json.load(f)
        return ZoneRaster(arr, meta["left"], meta["top"], meta["res"], meta.get("crs"), name=path)
    try:
        import rasterio  # type: ignore
    except ImportError as e:  # pragma: no cover - depends on the host image
        raise FileNotFoundError(
            f"cannot open raster '{path}': not a registered in-memory raster / .npy file and rasterio is "
            "not installed") from e
    if not os.path.isfile(path):  # pragma: no cover
        raise FileNotFoundError(path)
    with rasterio.open(path) as src:  # pragma: no cover
        return ZoneRaster(src.read(), src.bounds.left, src.bounds.top, abs(src.res[0]), str(src.crs), name=path)

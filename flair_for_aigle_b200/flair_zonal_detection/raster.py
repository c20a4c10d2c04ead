"""In-memory georeferenced raster: the part of ``rasterio.DatasetReader`` the hot path uses.

The reference opens rasters with rasterio (flair_zonal_detection/dataset.py:44-46,
inference.py:92-101, model_utils.py:11-16); this image has no rasterio/GDAL, and raster file
I/O is the "next" row of SURVEY.md section 8(f).  ``ZoneRaster`` carries what the path needs
-- pixels (C,H,W), bounds, resolution -- and can be built from a numpy array, a ``.npy`` file
(+ ``.json`` sidecar) or, when rasterio is installed, any GDAL-readable file.
"""
from __future__ import annotations

import json
import logging
import os
from collections import namedtuple
from typing import Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)

BoundingBox = namedtuple("BoundingBox", ["left", "bottom", "right", "top"])

_REGISTRY = {}


class ZoneRaster:
    def __init__(self, array: np.ndarray, left: float, top: float, res: float, crs: Optional[str] = None,
                 name: str = "<memory>"):
        if array.ndim == 2:
            array = array[None]
        self.array = array
        self.left, self.top, self.res_value, self.crs, self.name = float(left), float(top), float(res), crs, name
        self.pinned_tensor = None   # set by from_pinned(): the same pixels as a pinned torch tensor

    @classmethod
    def from_pinned(cls, tensor, left: float, top: float, res: float, crs: Optional[str] = None,
                    name: str = "<pinned>"):
        """Raster whose pixels live in page-locked host memory (torch uint8 tensor (C,H,W)), so the
        upload to HBM is a single asynchronous DMA."""
        r = cls(tensor.numpy(), left, top, res, crs, name)
        r.pinned_tensor = tensor
        return r

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

    def window_from_bounds(self, minx, miny, maxx, maxy):
        """``rasterio.windows.from_bounds(minx, miny, maxx, maxy, transform)`` (dataset.py:97): the FLOAT window
        (row_off, col_off, height, width) in this raster's pixels, through the inverse affine transform with the same
        arithmetic as affine.Affine.__invert__ / __mul__ (a = res, e = -res, b = d = 0).  Vectorised over arrays."""
        a, e, c, f = self.res_value, -self.res_value, self.left, self.top
        idet = 1.0 / (a * e)
        ra, re = e * idet, a * idet
        rb, rd = -0.0 * idet, -0.0 * idet
        rc, rf = -c * ra - f * rb, -c * rd - f * re
        minx, miny, maxx, maxy = (np.asarray(v, dtype=np.float64) for v in (minx, miny, maxx, maxy))
        cols = [x * ra + y * rb + rc for x, y in ((minx, maxy), (maxx, maxy), (maxx, miny), (minx, miny))]
        rows = [x * rd + y * re + rf for x, y in ((minx, maxy), (maxx, maxy), (maxx, miny), (minx, miny))]
        row0, row1 = np.minimum.reduce(rows), np.maximum.reduce(rows)
        col0, col1 = np.minimum.reduce(cols), np.maximum.reduce(cols)
        return row0, col0, np.maximum(row1 - row0, 0.0), np.maximum(col1 - col0, 0.0)

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
    """``rasterio.open(path)`` stand-in.  Accepts a ZoneRaster, a registered name, ``*.npy`` (with ``<path>.json`` =
    {left, top, res[, crs]}), an 8-bit north-up GeoTIFF (built-in reader, ``geotiff.py``) or, if rasterio is importable,
    any raster."""
    if isinstance(path, ZoneRaster):
        return path
    if path in _REGISTRY:
        return _REGISTRY[path]
    if isinstance(path, str) and path.endswith(".npy"):
        arr = np.load(path, mmap_mode="r")
        with open(path + ".json") as f:
            meta = json.load(f)
        return ZoneRaster(arr, meta["left"], meta["top"], meta["res"], meta.get("crs"), name=path)
    if isinstance(path, str) and path.lower().endswith((".tif", ".tiff")) and os.path.isfile(path):
        try:
            from .geotiff import read_geotiff
            arr, left, top, res, crs = read_geotiff(path)
            return ZoneRaster(arr, left, top, res, crs, name=path)
        except ValueError:
            pass            # not something the built-in reader handles: let rasterio try
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


_PINNED_POOL = {}
CLASSIC_TIFF_LIMIT = (1 << 32) - (1 << 24)      # bytes of pixel data a classic (32-bit offset) TIFF can hold, with headroom


class RasterSink:
    """Output raster of ``init_outputs`` (inference.py:157-208): a uint8 (count,H,W) array that lives
    on the GPU while tiles are written into it by the kernels, copied to the host once and stored
    on ``close()`` as a GeoTIFF (``geotiff.write_geotiff``: LZW for the single-band argmax raster like the
    reference's profile, Deflate planar for ``class_prob``), or as ``.npy`` when it exceeds classic TIFF;
    the georeferencing also goes to a ``.json`` sidecar."""

    def __init__(self, path: str, count: int, height: int, width: int, left: float, top: float, res: float,
                 crs=None, device=None):
        import torch
        self.name = path
        self.count, self.height, self.width = count, height, width
        self.left, self.top, self.res_value, self.crs = left, top, res, crs
        if device is None or torch.device(device).type != "cuda":
            raise RuntimeError("RasterSink needs a CUDA device: predictions are written by GPU kernels")
        self.device_array = torch.zeros((count, height, width), dtype=torch.uint8, device=device)
        self.host_array = None
        self.closed = False
        self._pinned = None

    def release(self) -> None:
        """Give the page-locked host buffer back for the next zone of the same shape (the numpy view returned by
        to_host() must not be used afterwards)."""
        if self._pinned is not None:
            _PINNED_POOL[tuple(self._pinned.shape)] = self._pinned
            self._pinned, self.host_array = None, None

    def pinned_buffer(self):
        """Page-locked host tensor of the raster's shape (reused across zones: page-locking 100 MB costs tens of ms)."""
        if self._pinned is None:
            import torch
            key = tuple(self.device_array.shape)
            pinned = _PINNED_POOL.pop(key, None)
            if pinned is None:
                pinned = torch.empty(self.device_array.shape, dtype=torch.uint8, pin_memory=True)
            self._pinned = pinned
        return self._pinned

    def mark_streamed(self) -> None:
        """The kernels' results were already read back into pinned_buffer() on a stream the current stream has joined
        (ZonalRunner.run_streamed): to_host() only synchronises."""
        self._streamed = True

    def to_host(self) -> np.ndarray:
        if self.host_array is None:
            import torch
            pinned = self.pinned_buffer()
            if not getattr(self, "_streamed", False):
                pinned.copy_(self.device_array, non_blocking=True)
            torch.cuda.current_stream(self.device_array.device).synchronize()
            self.host_array = pinned.numpy()
        return self.host_array

    write_files = True   # class-level switch: False = keep the result in host memory only

    def close(self) -> None:
        if self.closed:
            return
        arr = self.to_host()
        if not self.write_files:
            self.written_path = None
            self.closed = True
            return
        meta = {"left": self.left, "top": self.top, "res": self.res_value, "crs": self.crs,
                "count": self.count, "height": self.height, "width": self.width, "dtype": "uint8",
                "compress": "lzw"}
        from .geotiff import write_geotiff
        if arr.nbytes >= CLASSIC_TIFF_LIMIT:
            # classic TIFF addresses 4 GiB (a 60 000 x 60 000 class raster is 3.6 GB, its class_prob raster 68 GB): the
            # raster goes to <name>.npy + the .json sidecar instead -- said out loud, and ``written_path`` tells the caller
            written = os.path.splitext(self.name)[0] + ".npy"
            logger.warning(f"[!] {self.name}: {arr.nbytes / 2**30:.1f} GiB exceeds classic TIFF; writing {written} (+ .json "
                           "georeferencing) instead")
            np.save(written, arr)
        else:
            written = write_geotiff(self.name, arr, self.left, self.top, self.res_value, self.crs)   # errors propagate
        with open(written + ".json", "w") as f:
            json.dump(meta, f)
        self.written_path = written
        self.closed = True

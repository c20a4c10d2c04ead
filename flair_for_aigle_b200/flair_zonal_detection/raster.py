"""In-memory georeferenced raster: the part of ``rasterio.DatasetReader`` the hot path uses.

The reference opens rasters with rasterio (flair_zonal_detection/dataset.py:44-46,
inference.py:92-101, model_utils.py:11-16); this image has no rasterio/GDAL.  ``ZoneRaster`` carries what the path needs
-- pixels (C,H,W), bounds, resolution -- and can be built from a numpy array, a ``.npy`` file (+ ``.json`` sidecar), a
GeoTIFF / BigTIFF (``geotiff.read_geotiff``: decoded block-parallel by libfz_rasterio.so, on a GPU host straight into
page-locked memory so the upload to HBM needs no further copy), a JPEG-2000 file (``geotiff.read_jp2``) or, when rasterio
is installed, any GDAL-readable file.
"""
from __future__ import annotations

import json
import logging
import os
import threading
import weakref
from collections import namedtuple
from typing import Callable, NamedTuple, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)

BoundingBox = namedtuple("BoundingBox", ["left", "bottom", "right", "top"])

_REGISTRY = {}
_OPEN_FILES = weakref.WeakValueDictionary()     # (path, mtime, size) -> ZoneRaster, for as long as a caller holds it


class RowSource(NamedTuple):
    """A raster file that decodes any range of rows on request: ``read_rows(lo, hi, out)`` fills out (C, hi - lo, W).
    GeoTIFF: libfz_rasterio's block-parallel window read; JPEG 2000: OpenJPEG's decode area + worker threads."""
    read_rows: Callable[[int, int, np.ndarray], None]
    block_h: int            # rows of one block / tile row of the file: slabs and strips decode whole block rows once
    height: int
    width: int


class ProgressiveLoad:
    """A raster file being decoded on a background thread, bottom rows first, in slabs of whole block rows (each slab =
    one ``RowSource.read_rows`` into the destination's rows, itself parallel over the slab's blocks).  The zonal runner processes tile rows
    bottom-up and uploads only the rows the next batch needs (engine/zonal.py: run_streamed), so it calls
    ``wait_rows(lo)`` before each upload and the rest of the file decodes behind the forward pass."""

    def __init__(self, source: "RowSource", array: np.ndarray, tensor, slab_rows: int = 2048, row0: int = 0,
                 rows: Optional[int] = None):
        """Decodes file rows [row0, row0 + rows) into ``array`` (C, rows, W): the whole raster by default, one rank's row strip
        of it under multi-GPU sharding.  Row numbers of ``lo`` / ``wait_rows`` are relative to ``array``."""
        self.source, self.array, self.tensor = source, array, tensor
        self.row0 = int(row0)
        self.rows = int(source.height) - self.row0 if rows is None else int(rows)
        bh = max(1, int(source.block_h))
        self.slab = max(bh, (slab_rows // bh) * bh)
        self.lo = self.rows                              # rows [lo, rows) of the array are decoded
        self.error: Optional[BaseException] = None
        self._cv = threading.Condition()
        self._thread = threading.Thread(target=self._run, name="fz-raster-decode", daemon=True)
        self._thread.start()

    def _run(self) -> None:
        try:
            hi = self.row0 + self.rows                   # file rows; slab boundaries sit on whole block rows of the file
            while hi > self.row0:
                lo = max(self.row0, ((hi - 1) // self.slab) * self.slab)
                self.source.read_rows(lo, hi, self.array[:, lo - self.row0:hi - self.row0])
                with self._cv:
                    self.lo = lo - self.row0
                    self._cv.notify_all()
                hi = lo
        except BaseException as e:  # noqa: BLE001 -- handed to whoever waits
            with self._cv:
                self.error = e
                self._cv.notify_all()

    def wait_rows(self, lo: int, hi: Optional[int] = None) -> None:
        """Returns when every row >= lo is decoded (``hi`` is accepted for symmetry with the upload's row range)."""
        with self._cv:
            while self.lo > lo and self.error is None:
                self._cv.wait()
            if self.error is not None:
                raise self.error

    def wait_all(self) -> None:
        self.wait_rows(0)


class ZoneRaster:
    def __init__(self, array: np.ndarray, left: float, top: float, res: float, crs: Optional[str] = None,
                 name: str = "<memory>"):
        if array.ndim == 2:
            array = array[None]
        self._array, self._loader = array, None
        self._progressive_start, self._progress, self._source = None, None, None
        self._shape, self._dtype = tuple(array.shape), array.dtype
        self.left, self.top, self.res_value, self.crs, self.name = float(left), float(top), float(res), crs, name
        self.pinned_tensor = None   # set by from_pinned() / a lazy load: the same pixels as a pinned torch tensor

    @classmethod
    def from_pinned(cls, tensor, left: float, top: float, res: float, crs: Optional[str] = None,
                    name: str = "<pinned>"):
        """Raster whose pixels live in page-locked host memory (torch uint8 tensor (C,H,W)), so the
        upload to HBM is a single asynchronous DMA."""
        r = cls(tensor.numpy(), left, top, res, crs, name)
        r.pinned_tensor = tensor
        return r

    @classmethod
    def lazy(cls, shape, dtype, loader, left: float, top: float, res: float, crs: Optional[str] = None,
             name: str = "<file>", progressive=None, source: Optional["RowSource"] = None):
        """Raster file opened the way ``rasterio.open`` opens it: size, type and georeferencing now, pixels when somebody
        reads them.  ``loader() -> (array (C,H,W), pinned tensor or None)`` runs once, on the first access to ``array``:
        the geometry-only callers of the path (slicing.py:20-49, inference.py:76-132,157-208) never decode the file."""
        r = cls.__new__(cls)
        r._array, r._loader = None, loader
        r._shape, r._dtype = tuple(int(v) for v in shape), np.dtype(dtype)
        r.left, r.top, r.res_value, r.crs, r.name = float(left), float(top), float(res), crs, name
        r.pinned_tensor = None
        r._progressive_start, r._progress, r._source = progressive, None, source
        return r

    def begin_progressive(self) -> Optional["ProgressiveLoad"]:
        """Starts (or returns) the background decode of a lazily opened file: the destination array exists at once, rows
        become valid bottom-up (``ProgressiveLoad.wait_rows``).  None when the raster is already in memory or the file is not
        one the block decoder streams (JPEG 2000, Pillow fallback)."""
        if self._array is not None or self._progressive_start is None:
            return None
        if self._progress is None:
            self._progress = self._progressive_start()
        return self._progress

    def row_strip(self, r0: int, r1: int, name: Optional[str] = None) -> "ZoneRaster":
        """Rows [r0, r1) as a raster of their own (same columns, ``top`` moved down by r0 pixels): one rank's input strip of
        a zone sharded over GPUs (engine/strips.py).  In memory: a view, no copy.  A file the block decoder streams: only
        these rows are ever decoded (``fzio_read_window`` on the strip), lazily and progressively like the whole file."""
        r0, r1 = max(0, int(r0)), min(self.height, int(r1))
        if r1 <= r0:
            raise ValueError(f"{self.name}: empty row strip [{r0}, {r1})")
        top = self.top - r0 * self.res_value
        name = name or f"{self.name}#rows{r0}-{r1}"
        source = self._source
        if self._array is None and self._progress is None and source is not None:
            shape = (self._shape[0], r1 - r0, self._shape[2])

            def load():
                holder = {}
                arr = _pinned_array(shape, self._dtype, holder)
                source.read_rows(r0, r1, arr)
                return arr, holder.get("tensor")

            def progressive():
                holder = {}
                arr = _pinned_array(shape, self._dtype, holder)
                return ProgressiveLoad(source, arr, holder.get("tensor"), row0=r0, rows=r1 - r0)
            return ZoneRaster.lazy(shape, self._dtype, load, self.left, top, self.res_value, self.crs, name=name,
                                   progressive=progressive)
        arr = self.array                                  # in memory (or any other source: decoded once, then sliced)
        strip = ZoneRaster(arr[:, r0:r1], self.left, top, self.res_value, self.crs, name=name)
        if self.pinned_tensor is not None:
            strip.pinned_tensor = self.pinned_tensor[:, r0:r1]
        return strip

    @property
    def array(self) -> np.ndarray:
        if self._array is None:
            prog = self._progress
            if prog is not None:                          # a background decode is under way: its array, once complete
                prog.wait_all()
                arr, pinned = prog.array, prog.tensor
            else:
                arr, pinned = self._loader()
            if tuple(arr.shape) != self._shape:
                raise ValueError(f"{self.name}: decoded {tuple(arr.shape)}, the header said {self._shape}")
            self._array, self.pinned_tensor, self._loader, self._dtype = arr, pinned, None, arr.dtype
        return self._array

    @array.setter
    def array(self, value: np.ndarray) -> None:
        self._array, self._loader = value, None
        self._shape, self._dtype = tuple(value.shape), value.dtype

    @property
    def loaded(self) -> bool:
        return self._array is not None

    # -- rasterio.DatasetReader look-alikes ------------------------------------------------
    @property
    def count(self) -> int:
        return self._shape[0]

    @property
    def height(self) -> int:
        return self._shape[1]

    @property
    def width(self) -> int:
        return self._shape[2]

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
        return {"driver": "MEM", "dtype": str(self._dtype), "count": self.count, "height": self.height,
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
    {left, top, res[, crs]}), a north-up GeoTIFF / BigTIFF or JPEG-2000 file (``geotiff.py``; decoded into page-locked
    memory when the host has a CUDA device) or, if rasterio is importable, any raster."""
    if isinstance(path, ZoneRaster):
        return path
    if path in _REGISTRY:
        return _REGISTRY[path]
    if isinstance(path, str) and path.endswith(".npy"):
        arr = np.load(path, mmap_mode="r")
        with open(path + ".json") as f:
            meta = json.load(f)
        return ZoneRaster(arr, meta["left"], meta["top"], meta["res"], meta.get("crs"), name=path)
    if isinstance(path, str) and path.lower().endswith((".tif", ".tiff", ".jp2", ".j2k")) and os.path.isfile(path):
        st = os.stat(path)
        key = (os.path.abspath(path), st.st_mtime_ns, st.st_size)
        cached = _OPEN_FILES.get(key)
        if cached is not None:
            return cached                                 # same file, still held by somebody: one decode serves everyone
        from . import geotiff
        jp2 = path.lower().endswith((".jp2", ".j2k"))
        header, reader = (geotiff.jp2_header, geotiff.read_jp2) if jp2 else (geotiff.geotiff_header, geotiff.read_geotiff)
        try:
            shape, dtype, left, top, res, crs = header(path)
        except ValueError as e:
            not_ours = e                                  # not something the built-in readers handle: let rasterio try
        else:
            def load():
                holder = {}
                arr = reader(path, alloc=lambda shp, dt: _pinned_array(shp, dt, holder))[0]
                return arr, (holder.get("tensor") if arr is holder.get("array") else None)
            source = geotiff.row_source(path)             # None: the file is decoded in one go (Pillow paths)
            progressive = None
            if source is not None:
                def progressive():
                    holder = {}
                    arr = _pinned_array(shape, dtype, holder)
                    return ProgressiveLoad(source, arr, holder.get("tensor"))
            raster = ZoneRaster.lazy(shape, dtype, load, left, top, res, crs, name=path, progressive=progressive,
                                     source=source)       # row_strip() decodes row ranges of the source
            _OPEN_FILES[key] = raster
            return raster
        try:
            import rasterio  # type: ignore  # noqa: F401
        except ImportError:
            raise not_ours
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


def _pinned_array(shape, dtype, holder: dict) -> np.ndarray:
    """Array for a file reader to decode into: page-locked (a torch tensor's numpy view, kept in ``holder``) when the host
    has a CUDA device and the dtype is one the feeder uploads, plain numpy otherwise."""
    arr = None
    if np.dtype(dtype) in (np.dtype(np.uint8), np.dtype(np.float32)):
        try:
            import torch
            if torch.cuda.is_available():
                t = torch.empty(tuple(shape), dtype=torch.uint8 if np.dtype(dtype) == np.uint8 else torch.float32,
                                pin_memory=True)
                holder["tensor"] = t
                arr = t.numpy()
        except RuntimeError:  # pragma: no cover - page-locking can fail on exotic hosts
            holder.pop("tensor", None)
            arr = None
    if arr is None:
        arr = np.empty(tuple(shape), dtype)
    holder["array"] = arr
    return arr


_PINNED_POOL = {}


class RasterSink:
    """Output raster of ``init_outputs`` (inference.py:157-208): a uint8 (count,H,W) array that lives
    on the GPU while tiles are written into it by the kernels, copied to the host once and stored
    on ``close()`` as a tiled LZW GeoTIFF like the reference's profile (``geotiff.write_geotiff``: every 512 x 512 block
    encoded on its own host core by libfz_rasterio.so; ``class_prob`` = one plane per class; BigTIFF when the file
    outgrows 32-bit offsets); the georeferencing also goes to a ``.json`` sidecar."""

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
        written = write_geotiff(self.name, arr, self.left, self.top, self.res_value, self.crs)   # errors propagate
        with open(written + ".json", "w") as f:
            json.dump(meta, f)
        self.written_path = written
        self.closed = True

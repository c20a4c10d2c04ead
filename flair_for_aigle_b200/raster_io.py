"""ctypes binding of libfz_rasterio.so (C ABI declared in include/flair_zonal_rasterio.h): block-parallel TIFF / BigTIFF /
GeoTIFF reading and writing on the host cores, the file I/O at the two ends of the zonal path.

numpy arrays in, numpy arrays out; no torch, no CUDA, no GDAL.  A missing library raises ``RasterIOError`` (there is no
pure-Python fallback behind these calls).
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import NamedTuple, Optional, Sequence

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "_native" / "libfz_rasterio.so"
_lib: Optional[ctypes.CDLL] = None

ABI_VERSION = 1
COMP_NONE, COMP_LZW, COMP_DEFLATE = 1, 5, 8
FMT_UINT, FMT_INT, FMT_FLOAT = 1, 2, 3
OVR_NEAREST, OVR_MODE = 0, 1
_COMPRESSION = {"none": COMP_NONE, "lzw": COMP_LZW, "deflate": COMP_DEFLATE}


class RasterIOError(RuntimeError):
    pass


class _Info(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int64), ("height", ctypes.c_int64)] + \
               [(k, ctypes.c_int32) for k in ("count", "bits", "sample_format", "compression", "predictor", "planar", "tiled",
                                              "block_w", "block_h", "bigtiff", "big_endian", "overviews", "has_georef", "epsg",
                                              "geographic")] + \
               [("_pad", ctypes.c_int32)] + \
               [(k, ctypes.c_double) for k in ("left", "top", "res_x", "res_y")]


class _WriteOpts(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("block", "compression", "predictor", "deflate_level", "pixel_interleave",
                                              "overviews", "overview_resampling", "cog", "bigtiff", "threads",
                                              "sample_format", "bits", "epsg", "geographic", "has_georef", "reserved")] + \
               [(k, ctypes.c_double) for k in ("left", "top", "res")]


class TiffInfo(NamedTuple):
    width: int
    height: int
    count: int
    dtype: np.dtype
    compression: int
    predictor: int
    planar: int
    tiled: bool
    block_w: int
    block_h: int
    bigtiff: bool
    overviews: int
    has_georef: bool
    epsg: int
    geographic: bool
    left: float
    top: float
    res_x: float
    res_y: float

    @property
    def crs(self) -> Optional[str]:
        return f"EPSG:{self.epsg}" if self.epsg else None


_vp, _i, _i64, _cp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_char_p
_SIGNATURES = {
    "fzio_abi_version": ([], _i),
    "fzio_last_error": ([], _cp),
    "fzio_tiff_info": ([_cp, _i, ctypes.POINTER(_Info)], _i),
    "fzio_read_window": ([_cp, _i, _i64, _i64, _i64, _i64, _vp, _i, _vp, _i64, _i64, _i], _i),
    "fzio_write_geotiff": ([_cp, _vp, _i, _i64, _i64, _i64, _i64, ctypes.POINTER(_WriteOpts)], _i),
    "fzio_convert_to_cog": ([_cp, _cp, _i], _i),
    "fzio_lzw_bound": ([_i64], _i64),
    "fzio_lzw_encode": ([_vp, _i64, _vp, _i64], _i64),
    "fzio_lzw_decode": ([_vp, _i64, _vp, _i64], _i64),
}


def exported_symbols():
    return list(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RasterIOError(f"{_LIB_PATH} is missing: run `python -m flair_for_aigle_b200.build` "
                                "(the raster I/O library has no Python fallback)")
        handle = ctypes.CDLL(str(_LIB_PATH))
        for name, (args, res) in _SIGNATURES.items():
            if not hasattr(handle, name):
                raise RasterIOError(f"{_LIB_PATH}: symbol {name} is missing (stale build?)")
            fn = getattr(handle, name)
            fn.argtypes, fn.restype = args, res
        if handle.fzio_abi_version() != ABI_VERSION:
            raise RasterIOError(f"{_LIB_PATH}: ABI version {handle.fzio_abi_version()}, binding expects {ABI_VERSION}")
        _lib = handle
    return _lib


def host_threads(threads: int = 0) -> int:
    """Worker threads for one call: an explicit positive request wins; otherwise ``FZ_IO_THREADS``; otherwise, in a
    one-process-per-GPU job (torchrun sets LOCAL_WORLD_SIZE), this process's share of the box's cores -- eight ranks each
    decoding their strip of a zone must not start eight times the core count; otherwise 0 = all cores."""
    if threads and threads > 0:
        return int(threads)
    env = os.environ.get("FZ_IO_THREADS", "")
    if env.isdigit() and int(env) > 0:
        return int(env)
    local_world = os.environ.get("LOCAL_WORLD_SIZE", "")
    if local_world.isdigit() and int(local_world) > 1:
        return max(1, (os.cpu_count() or 1) // int(local_world))
    return 0


def _check(rc: int, what: str) -> None:
    if rc < 0:
        raise RasterIOError(f"{what}: {lib().fzio_last_error().decode(errors='replace')}")


def _dtype(bits: int, fmt: int) -> np.dtype:
    try:
        return np.dtype({(8, 1): "u1", (16, 1): "u2", (32, 1): "u4", (8, 2): "i1", (16, 2): "i2", (32, 2): "i4",
                         (32, 3): "f4"}[(bits, fmt)])
    except KeyError:
        raise RasterIOError(f"{bits}-bit samples of format {fmt} are not supported") from None


def _format(dtype) -> tuple:
    dt = np.dtype(dtype)
    table = {"u1": (8, 1), "u2": (16, 1), "u4": (32, 1), "i1": (8, 2), "i2": (16, 2), "i4": (32, 2), "f4": (32, 3)}
    key = dt.str.lstrip("<|=")
    if key not in table:
        raise RasterIOError(f"dtype {dt} cannot be written (uint8/16/32, int8/16/32, float32)")
    return table[key]


def tiff_info(path: str, level: int = 0) -> TiffInfo:
    """What ``rasterio.open(path)`` exposes to the path (inference.py:92-101): size, band count, dtype, block layout, bounds
    origin, resolution, EPSG code.  ``level`` k > 0 = the k-th overview."""
    info = _Info()
    _check(lib().fzio_tiff_info(os.fsencode(path), int(level), ctypes.byref(info)), "tiff_info")
    return TiffInfo(info.width, info.height, info.count, _dtype(info.bits, info.sample_format), info.compression,
                    info.predictor, info.planar, bool(info.tiled), info.block_w, info.block_h, bool(info.bigtiff),
                    info.overviews, bool(info.has_georef), info.epsg, bool(info.geographic), info.left, info.top,
                    info.res_x, info.res_y)


def read_window(path: str, row0: int, col0: int, height: int, width: int, bands: Optional[Sequence[int]] = None,
                out: Optional[np.ndarray] = None, level: int = 0, threads: int = 0, info: Optional[TiffInfo] = None) -> np.ndarray:
    """``src.read(indexes=bands, window=Window(col0, row0, width, height), boundless=True, fill_value=0)``
    (dataset.py:108-115) -> (n_bands, height, width) in the file's dtype.  ``out``: a C-contiguous (or row-strided) array
    to decode into -- e.g. the numpy view of a page-locked torch tensor, so the decoded raster is upload-ready."""
    info = info or tiff_info(path, level)
    n = info.count if bands is None else len(bands)
    if out is None:
        out = np.empty((n, height, width), info.dtype)
    if out.dtype != info.dtype or out.shape != (n, height, width):
        raise RasterIOError(f"read_window: out is {out.dtype}{out.shape}, the window needs {info.dtype}{(n, height, width)}")
    if out.size and out.strides[2] != out.itemsize:
        raise RasterIOError("read_window: out rows must be contiguous")
    band_arr = None if bands is None else (ctypes.c_int32 * n)(*[int(b) for b in bands])
    _check(lib().fzio_read_window(os.fsencode(path), int(level), int(row0), int(col0), int(height), int(width), band_arr, n,
                                  out.ctypes.data, out.strides[0] if out.size else 0, out.strides[1] if out.size else 0,
                                  host_threads(threads)), "read_window")
    return out


def read_raster(path: str, bands: Optional[Sequence[int]] = None, out: Optional[np.ndarray] = None, level: int = 0,
                threads: int = 0):
    """The whole raster, decoded block-parallel -> (array (count, H, W), TiffInfo)."""
    info = tiff_info(path, level)
    return read_window(path, 0, 0, info.height, info.width, bands, out, level, threads, info), info


def write_geotiff(path: str, arr: np.ndarray, left: Optional[float] = None, top: Optional[float] = None,
                  res: Optional[float] = None, epsg: int = 0, geographic: bool = False, compression: str = "lzw",
                  block: int = 512, predictor: int = 1, deflate_level: int = 6, pixel_interleave: bool = False,
                  overviews: int = 0, overview_resampling: str = "nearest", cog: bool = False, bigtiff: int = 0,
                  threads: int = 0) -> str:
    """arr (count, H, W) or (H, W) -> tiled GeoTIFF (BigTIFF when classic offsets cannot address it), every block compressed
    on its own host thread.  Defaults = the reference's output profile (``compress='lzw'``, inference.py:182-203)."""
    if arr.ndim == 2:
        arr = arr[None]
    if arr.ndim != 3:
        raise RasterIOError(f"write_geotiff: array of shape {arr.shape}")
    if arr.strides[2] != arr.itemsize or arr.strides[1] < 0 or arr.strides[0] < 0:
        arr = np.ascontiguousarray(arr)
    bits, fmt = _format(arr.dtype)
    if compression not in _COMPRESSION:
        raise RasterIOError(f"compression '{compression}' (none, lzw, deflate)")
    if overview_resampling not in ("nearest", "mode"):
        raise RasterIOError(f"overview_resampling '{overview_resampling}' (nearest, mode)")
    o = _WriteOpts()
    o.block, o.compression, o.predictor, o.deflate_level = int(block), _COMPRESSION[compression], int(predictor), int(deflate_level)
    o.pixel_interleave, o.overviews, o.cog, o.bigtiff = int(pixel_interleave), int(overviews), int(cog), int(bigtiff)
    o.threads = host_threads(threads)
    o.overview_resampling = OVR_MODE if overview_resampling == "mode" else OVR_NEAREST
    o.sample_format, o.bits = fmt, bits
    if left is not None and top is not None and res is not None:
        o.has_georef, o.left, o.top, o.res = 1, float(left), float(top), float(res)
        o.epsg, o.geographic = int(epsg or 0), int(bool(geographic))
    count, h, w = arr.shape
    _check(lib().fzio_write_geotiff(os.fsencode(path), arr.ctypes.data, count, h, w, arr.strides[0], arr.strides[1],
                                    ctypes.byref(o)), "write_geotiff")
    return path


def convert_to_cog(src_path: str, dst_path: str, threads: int = 0) -> str:
    """postprocess.py:33-52: GeoTIFF -> COG (LZW, 512 x 512 blocks, nearest overviews down to one block, IFDs first)."""
    _check(lib().fzio_convert_to_cog(os.fsencode(src_path), os.fsencode(dst_path), host_threads(threads)), "convert_to_cog")
    return dst_path


def lzw_encode(data: bytes) -> bytes:
    src = np.frombuffer(data, np.uint8)
    cap = lib().fzio_lzw_bound(src.size)
    dst = np.empty(cap, np.uint8)
    n = lib().fzio_lzw_encode(src.ctypes.data if src.size else None, src.size, dst.ctypes.data, cap)
    _check(n, "lzw_encode")
    return dst[:n].tobytes()


def lzw_decode(data: bytes, size: int) -> bytes:
    src = np.frombuffer(data, np.uint8)
    dst = np.empty(size, np.uint8)
    n = lib().fzio_lzw_decode(src.ctypes.data if src.size else None, src.size, dst.ctypes.data if size else None, size)
    _check(n, "lzw_decode")
    return dst[:n].tobytes()

"""JPEG 2000 decoding through the OpenJPEG library itself (ctypes), for the reference's product inputs (``*.jp2`` orthophotos,
flair_zonal_detection/inference.py:60; read there by rasterio / GDAL one window per tile, dataset.py:108-115).

Pillow ships libopenjp2 but drives it one way only: the whole image, one thread.  Binding the library directly gives what the
zonal path needs from a decoder: ROW-WINDOW decodes (``opj_set_decode_area``: a rank's strip of a zone shared over GPUs; slabs
decoded bottom-up behind the upload, raster.ProgressiveLoad) and OpenJPEG's own worker threads (``opj_codec_set_threads``).
The library is the one inside Pillow's wheel (``pillow.libs/libopenjp2*.so``); if it cannot be found, is not a 2.x release,
or a header looks implausible, ``OpenJPEGUnavailable`` is raised and ``geotiff.read_jp2`` falls back to Pillow.

Only what is needed is bound: 8-bit unsigned components without subsampling (orthophotos); anything else raises.
"""
from __future__ import annotations

import ctypes
import glob
import os
import threading
from typing import NamedTuple, Optional

import numpy as np

OPJ_CODEC_J2K, OPJ_CODEC_JP2 = 0, 2


class OpenJPEGUnavailable(RuntimeError):
    pass


class OpenJPEGError(RuntimeError):
    pass


class _Comp(ctypes.Structure):          # opj_image_comp_t (openjpeg.h, 2.x; ``bpp`` is deprecated but still in the struct)
    _fields_ = [("dx", ctypes.c_uint32), ("dy", ctypes.c_uint32), ("w", ctypes.c_uint32), ("h", ctypes.c_uint32),
                ("x0", ctypes.c_uint32), ("y0", ctypes.c_uint32), ("prec", ctypes.c_uint32), ("bpp", ctypes.c_uint32),
                ("sgnd", ctypes.c_uint32), ("resno_decoded", ctypes.c_uint32), ("factor", ctypes.c_uint32),
                ("data", ctypes.POINTER(ctypes.c_int32)), ("alpha", ctypes.c_uint16)]


class _Image(ctypes.Structure):         # opj_image_t
    _fields_ = [("x0", ctypes.c_uint32), ("y0", ctypes.c_uint32), ("x1", ctypes.c_uint32), ("y1", ctypes.c_uint32),
                ("numcomps", ctypes.c_uint32), ("color_space", ctypes.c_int), ("comps", ctypes.POINTER(_Comp)),
                ("icc_profile_buf", ctypes.c_void_p), ("icc_profile_len", ctypes.c_uint32)]


class _CstrInfoHead(ctypes.Structure):  # the first members of opj_codestream_info_v2_t: the tile grid
    _fields_ = [("tx0", ctypes.c_uint32), ("ty0", ctypes.c_uint32), ("tdx", ctypes.c_uint32), ("tdy", ctypes.c_uint32),
                ("tw", ctypes.c_uint32), ("th", ctypes.c_uint32), ("nbcomps", ctypes.c_uint32)]


_MSG_CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.c_void_p)
_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        try:
            import PIL
        except ImportError as e:  # pragma: no cover
            raise OpenJPEGUnavailable("Pillow (whose wheel carries libopenjp2) is not installed") from e
        base = os.path.dirname(os.path.dirname(os.path.abspath(PIL.__file__)))
        cands = sorted(glob.glob(os.path.join(base, "pillow.libs", "libopenjp2*.so*")) +
                       glob.glob(os.path.join(base, "Pillow.libs", "libopenjp2*.so*")))
        if not cands:
            raise OpenJPEGUnavailable("no libopenjp2 inside Pillow's wheel (pillow.libs/)")
        try:
            h = ctypes.CDLL(cands[0])
        except OSError as e:  # pragma: no cover
            raise OpenJPEGUnavailable(f"{cands[0]}: {e}") from e
        vp, cp, i32 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32
        sigs = {"opj_version": ([], cp), "opj_has_thread_support": ([], ctypes.c_int),
                "opj_stream_create_default_file_stream": ([cp, ctypes.c_int], vp), "opj_stream_destroy": ([vp], None),
                "opj_create_decompress": ([ctypes.c_int], vp), "opj_destroy_codec": ([vp], None),
                "opj_set_default_decoder_parameters": ([vp], None), "opj_setup_decoder": ([vp, vp], ctypes.c_int),
                "opj_codec_set_threads": ([vp, ctypes.c_int], ctypes.c_int),
                "opj_set_error_handler": ([vp, _MSG_CB, vp], ctypes.c_int),
                "opj_set_warning_handler": ([vp, _MSG_CB, vp], ctypes.c_int),
                "opj_read_header": ([vp, vp, ctypes.POINTER(ctypes.POINTER(_Image))], ctypes.c_int),
                "opj_set_decode_area": ([vp, ctypes.POINTER(_Image), i32, i32, i32, i32], ctypes.c_int),
                "opj_decode": ([vp, vp, ctypes.POINTER(_Image)], ctypes.c_int),
                "opj_end_decompress": ([vp, vp], ctypes.c_int), "opj_image_destroy": ([ctypes.POINTER(_Image)], None),
                "opj_get_cstr_info": ([vp], ctypes.POINTER(_CstrInfoHead)),
                "opj_destroy_cstr_info": ([ctypes.POINTER(ctypes.POINTER(_CstrInfoHead))], None)}
        for name, (args, res) in sigs.items():
            try:
                fn = getattr(h, name)
            except AttributeError as e:
                raise OpenJPEGUnavailable(f"{cands[0]}: no symbol {name}") from e
            fn.argtypes, fn.restype = args, res
        version = h.opj_version().decode()
        if not version.startswith("2."):
            raise OpenJPEGUnavailable(f"libopenjp2 {version}: the structures bound here are those of the 2.x releases")
        _lib = h
        return _lib


class JP2Info(NamedTuple):
    width: int
    height: int
    count: int
    tile_w: int
    tile_h: int
    threads_supported: bool


class _Decoder:
    """One open codestream: stream + codec + the image header; closed on exit."""

    def __init__(self, path: str, threads: int = 0):
        self.h = lib()
        with open(path, "rb") as f:
            magic = f.read(12)
        if magic[4:8] == b"jP  ":
            fmt = OPJ_CODEC_JP2
        elif magic[:4] == b"\xff\x4f\xff\x51":
            fmt = OPJ_CODEC_J2K
        else:
            raise OpenJPEGError(f"{path}: neither a JP2 file nor a JPEG 2000 codestream")
        self.path, self.errors = path, []
        self._cb = _MSG_CB(lambda msg, _: self.errors.append((msg or b"").decode(errors="replace").strip()))
        self._quiet = _MSG_CB(lambda msg, _: None)
        self.stream = self.codec = None
        self.image = ctypes.POINTER(_Image)()
        try:
            self.stream = self.h.opj_stream_create_default_file_stream(os.fsencode(path), 1)
            if not self.stream:
                raise OpenJPEGError(f"{path}: cannot open")
            self.codec = self.h.opj_create_decompress(fmt)
            self.h.opj_set_error_handler(self.codec, self._cb, None)
            self.h.opj_set_warning_handler(self.codec, self._quiet, None)
            params = ctypes.create_string_buffer(32768)          # opj_dparameters_t (~8.3 KB), defaults only
            self.h.opj_set_default_decoder_parameters(params)
            if not self.h.opj_setup_decoder(self.codec, params):
                raise OpenJPEGError(f"{path}: opj_setup_decoder failed")
            if threads != 1 and self.h.opj_has_thread_support():
                from .raster_io import host_threads
                n = host_threads(threads) or (os.cpu_count() or 1)      # this process's share of the cores under torchrun
                if n > 1:
                    self.h.opj_codec_set_threads(self.codec, n)
            if not self.h.opj_read_header(self.stream, self.codec, ctypes.byref(self.image)) or not self.image:
                raise OpenJPEGError(f"{path}: {'; '.join(self.errors) or 'opj_read_header failed'}")
            im = self.image.contents
            self.x0, self.y0 = int(im.x0), int(im.y0)
            self.width, self.height, self.count = int(im.x1) - self.x0, int(im.y1) - self.y0, int(im.numcomps)
            if not (0 < self.width <= 1 << 20 and 0 < self.height <= 1 << 20 and 0 < self.count <= 64):
                raise OpenJPEGUnavailable(f"{path}: implausible header {self.width} x {self.height} x {self.count} "
                                          "(structure layout mismatch?)")
            for c in range(self.count):
                comp = im.comps[c]
                if (comp.dx, comp.dy) != (1, 1) or comp.prec != 8 or comp.sgnd:
                    raise OpenJPEGError(f"{path}: component {c} is {comp.prec}-bit{' signed' if comp.sgnd else ''}, subsampling "
                                        f"{comp.dx} x {comp.dy}; only 8-bit unsigned, unsubsampled components are decoded")
        except BaseException:
            self.close()
            raise

    def tile_grid(self):
        p = self.h.opj_get_cstr_info(self.codec)
        if not p:
            return self.width, self.height
        try:
            t = p.contents
            tdx, tdy = int(t.tdx), int(t.tdy)
        finally:
            self.h.opj_destroy_cstr_info(ctypes.byref(p))
        if not (0 < tdx <= 1 << 24 and 0 < tdy <= 1 << 24):
            return self.width, self.height
        return min(tdx, self.width), min(tdy, self.height)

    def decode_rows(self, row0: int, row1: int, out: np.ndarray) -> None:
        """Rows [row0, row1) of every component -> out (count, row1 - row0, width) uint8."""
        if not (0 <= row0 < row1 <= self.height):
            raise OpenJPEGError(f"{self.path}: rows [{row0}, {row1}) outside 0..{self.height}")
        if (row0, row1) != (0, self.height):
            if not self.h.opj_set_decode_area(self.codec, self.image, self.x0, self.y0 + row0, self.x0 + self.width, self.y0 + row1):
                raise OpenJPEGError(f"{self.path}: {'; '.join(self.errors) or 'opj_set_decode_area failed'}")
        if not self.h.opj_decode(self.codec, self.stream, self.image):
            raise OpenJPEGError(f"{self.path}: {'; '.join(self.errors) or 'opj_decode failed'}")
        im = self.image.contents
        rows = row1 - row0
        for c in range(self.count):
            comp = im.comps[c]
            if (int(comp.w), int(comp.h)) != (self.width, rows) or not comp.data:
                raise OpenJPEGError(f"{self.path}: component {c} decoded as {comp.w} x {comp.h}, expected {self.width} x {rows}")
            plane = np.ctypeslib.as_array(comp.data, shape=(rows, self.width))
            np.copyto(out[c], plane, casting="unsafe")            # int32 samples 0..255 -> uint8, into the caller's buffer

    def close(self) -> None:
        if self.image:
            self.h.opj_image_destroy(self.image)
            self.image = ctypes.POINTER(_Image)()
        if self.codec:
            self.h.opj_destroy_codec(self.codec)
            self.codec = None
        if self.stream:
            self.h.opj_stream_destroy(self.stream)
            self.stream = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def info(path: str) -> JP2Info:
    with _Decoder(path, threads=1) as d:
        tw, th = d.tile_grid()
        return JP2Info(d.width, d.height, d.count, tw, th, bool(d.h.opj_has_thread_support()))


def read_rows(path: str, row0: int = 0, row1: Optional[int] = None, out: Optional[np.ndarray] = None,
              threads: int = 0) -> np.ndarray:
    """Rows [row0, row1) of the image (all of it by default) -> (count, rows, width) uint8.  One decoder per call (an OpenJPEG
    codec decodes one area once); ``threads``: OpenJPEG worker threads, 0 = all cores."""
    with _Decoder(path, threads=threads) as d:
        row1 = d.height if row1 is None else row1
        shape = (d.count, row1 - row0, d.width)
        if out is None:
            out = np.empty(shape, np.uint8)
        if out.shape != shape or out.dtype != np.uint8:
            raise OpenJPEGError(f"read_rows: out is {out.dtype}{out.shape}, the rows need uint8{shape}")
        d.decode_rows(row0, row1, out)
        return out

"""Raster files at the two ends of the zonal path, on the host (rasterio/GDAL are not required).  Replaces, for north-up
rasters:

  rasterio.open(path).read() / .bounds / .res / .crs      flair_zonal_detection/dataset.py:89-117, inference.py:76-132
  rasterio.open(path, 'w', **profile).write(...)          flair_zonal_detection/inference.py:157-208,343-352

GeoTIFF / BigTIFF in and out go through ``libfz_rasterio.so`` (``..raster_io``; include/flair_zonal_rasterio.h): every block
is decoded / LZW-encoded on its own host core, straight into / out of the (page-locked) array the GPU transfer uses.
Outputs are tiled LZW GeoTIFFs (the reference's ``compress='lzw'`` profile; ``class_prob`` = one plane per class), BigTIFF
when the file outgrows 32-bit offsets (a 60 000 x 60 000 zone's ``class_prob`` raster).  Georeferencing is carried by
ModelPixelScaleTag (33550), ModelTiepointTag (33922) and a GeoKeyDirectoryTag (34735) holding the EPSG code, which is what
GDAL writes for such rasters.  TIFF codecs the library does not implement (JPEG-in-TIFF ...) fall back to libtiff through
Pillow, one core.  JPEG-2000 inputs -- what the reference's product script globs for (inference.py:60) -- are decoded by
OpenJPEG through Pillow and georeferenced from the GeoJP2 box, the GMLJP2 box or a world file (``read_jp2``).
"""
from __future__ import annotations

import os
import re
import struct
from typing import Callable, Optional, Tuple

import numpy as np

from .. import raster_io

TAG_PIXEL_SCALE, TAG_TIEPOINT, TAG_GEOKEYS = 33550, 33922, 34735


def _epsg(crs: Optional[str]) -> Optional[int]:
    """EPSG code of a CRS given as ``EPSG:<n>`` / ``urn:ogc:def:crs:EPSG::<n>`` / a bare number, or as WKT whose OUTERMOST
    authority is EPSG (``AUTHORITY["EPSG","2154"]]`` / ``ID["EPSG",2154]]`` closing the string).  Anything else (a WKT without
    an authority, a PROJ string) has no code: no digits are guessed out of names such as ``GRS 1980``."""
    if crs is None:
        return None
    text = str(crs).strip()
    m = re.fullmatch(r"(?:EPSG|epsg)\s*:\s*(\d+)", text) or re.fullmatch(r"urn:ogc:def:crs:EPSG:[^:]*:(\d+)", text, flags=re.I) \
        or re.fullmatch(r"(\d{4,6})", text)
    if m:
        return int(m.group(1))
    m = re.search(r"(?:AUTHORITY\s*\[\s*\"EPSG\"\s*,\s*\"?(\d+)\"?\s*\]|ID\s*\[\s*\"EPSG\"\s*,\s*(\d+)\s*\])\s*\]\s*$", text)
    if m:
        return int(m.group(1) or m.group(2))
    return None


def _is_geographic(crs: Optional[str], epsg: Optional[int]) -> bool:
    """Geographic (lon/lat) CRS: a WKT starting with GEOGCS / GEOGCRS / GEODCRS, or one of the usual geographic EPSG codes
    (4000-4999: WGS 84 = 4326, RGF93 = 4171, ETRS89 = 4258 ...)."""
    text = str(crs or "").lstrip().upper()
    if text.startswith(("GEOGCS", "GEOGCRS", "GEODCRS", "GEOGRAPHICCRS")):
        return True
    if text.startswith(("PROJCS", "PROJCRS", "PROJECTEDCRS")):
        return False
    return epsg is not None and 4000 <= epsg <= 4999


def _geokeys(crs: Optional[str]):
    """GeoKeyDirectory version 1.1.0: GTModelType (1 projected / 2 geographic), GTRasterType = PixelIsArea, and the EPSG code
    under ProjectedCSTypeGeoKey (3072) or GeographicTypeGeoKey (2048)."""
    epsg = _epsg(crs)
    geographic = _is_geographic(crs, epsg)
    keys = [(1024, 0, 1, 2 if geographic else 1), (1025, 0, 1, 1)]
    if epsg is not None:
        keys.append((2048 if geographic else 3072, 0, 1, epsg))
    keys.sort()
    flat = [1, 1, 0, len(keys)]
    for k in keys:
        flat.extend(k)
    return tuple(flat)


def write_geotiff(path: str, arr: np.ndarray, left: float, top: float, res: float, crs: Optional[str] = None) -> str:
    """arr (count, H, W) uint8 (class rasters) or any dtype the library writes -> tiled LZW GeoTIFF, every 512 x 512 block
    encoded on its own host core; BigTIFF when needed.  Returns the path written."""
    assert arr.ndim == 3, arr.shape
    epsg = _epsg(crs)
    try:
        return raster_io.write_geotiff(path, arr, left, top, res, epsg=epsg or 0, geographic=_is_geographic(crs, epsg),
                                       compression="lzw")
    except raster_io.RasterIOError as e:
        raise ValueError(str(e)) from e


Allocator = Callable[[tuple, np.dtype], np.ndarray]


def _square_res(path: str, rx: float, ry: float) -> float:
    if abs(rx - ry) > 1e-9 * max(rx, ry):
        raise ValueError(f"{path}: non-square pixels ({rx} x {ry}) are not supported")
    return rx


def _tiff_info(path: str):
    """-> (TiffInfo, left, top, res, crs).  Georeferencing from the GeoTIFF tags or, for a plain TIFF, from a world file
    (.tfw) or a MapInfo .tab registration next to it (GDAL's fallbacks; how IGN delivers some orthophoto tiles)."""
    try:
        info = raster_io.tiff_info(path)
    except raster_io.RasterIOError as e:
        raise ValueError(str(e)) from e
    if info.has_georef:
        return info, info.left, info.top, _square_res(path, info.res_x, info.res_y), info.crs
    side = _world_file(path) or _tab_file(path)
    if side is None:
        raise ValueError(f"{path}: no GeoTIFF georeferencing (ModelPixelScale / ModelTiepoint tags), world file or .tab")
    left, top, sx, sy, crs = side
    return info, left, top, _square_res(path, sx, sy), crs


def geotiff_header(path: str):
    """-> (shape (count, H, W), dtype, left, top, res, crs) from the directory alone: no pixel is decoded."""
    info, left, top, res, crs = _tiff_info(path)
    return (info.count, info.height, info.width), info.dtype, left, top, res, crs


def row_source(path: str):
    """``raster.RowSource`` of a file whose rows can be decoded range by range -- so it can be decoded slab by slab behind the
    upload, and strip by strip over several GPUs -- or None (the file is then decoded in one go).  GeoTIFF: when
    libfz_rasterio decodes the file itself; JPEG 2000: when the OpenJPEG library can be driven directly."""
    from .raster import RowSource
    if path.lower().endswith((".jp2", ".j2k")):
        from .. import openjpeg
        if os.environ.get("FZ_JP2_DECODER", "").lower() == "pillow":      # switch the direct OpenJPEG binding off
            return None
        try:
            j = openjpeg.info(path)
        except Exception:  # noqa: BLE001 -- no library, a file it does not decode, a binding problem: Pillow's path takes over
            return None
        def read_jp2_rows(lo: int, hi: int, out: np.ndarray) -> None:
            openjpeg.read_rows(path, lo, hi, out=out)
        return RowSource(read_jp2_rows, j.tile_h, j.height, j.width)
    try:
        info = raster_io.tiff_info(path)
    except raster_io.RasterIOError:
        return None
    ok = (info.compression in (raster_io.COMP_NONE, raster_io.COMP_LZW, raster_io.COMP_DEFLATE) and info.predictor in (1, 2, 3)
          and info.planar in (1, 2))
    if not ok:
        return None
    def read_tiff_rows(lo: int, hi: int, out: np.ndarray) -> None:
        raster_io.read_window(path, lo, 0, hi - lo, info.width, out=out, info=info)
    return RowSource(read_tiff_rows, info.block_h, info.height, info.width)


def read_geotiff(path: str, alloc: Optional[Allocator] = None) -> Tuple[np.ndarray, float, float, float, Optional[str]]:
    """-> (array (count, H, W), left, top, res, crs).  ``alloc(shape, dtype)`` supplies the array to decode into (e.g. the
    numpy view of a page-locked tensor).  Strips or tiles, classic or BigTIFF, none / LZW / Deflate, predictor 2, 8 / 16 /
    32-bit samples: decoded block-parallel by libfz_rasterio; any other libtiff codec: Pillow."""
    info, left, top, res, crs = _tiff_info(path)
    shape = (info.count, info.height, info.width)
    try:
        out = alloc(shape, info.dtype) if alloc is not None else None
        arr = raster_io.read_window(path, 0, 0, info.height, info.width, out=out, info=info)
    except raster_io.RasterIOError as e:
        if "is not supported" not in str(e):
            raise ValueError(str(e)) from e
        arr = _read_with_pillow(path)
    return arr, left, top, res, crs


def _read_with_pillow(path: str) -> np.ndarray:
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        a = np.asarray(im)
    if a.dtype != np.uint8:
        raise ValueError(f"{path}: only 8-bit rasters are supported through the Pillow fallback (got {a.dtype})")
    return a[None] if a.ndim == 2 else np.ascontiguousarray(a.transpose(2, 0, 1))


# ------------------------------------------------------------------------------------------------------------ JPEG 2000
_GEOJP2_UUID = bytes.fromhex("b14bf8bd083d4b43a5ae8cd7d5a6ce03")


def _jp2_boxes(data: bytes, start: int = 0, end: Optional[int] = None):
    """(type, payload_start, payload_end) of the boxes in data[start:end] (ISO 15444-1 Annex I)."""
    end = len(data) if end is None else end
    pos = start
    while pos + 8 <= end:
        size, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = pos + 8
        if size == 1:
            if pos + 16 > end:
                return
            size = struct.unpack(">Q", data[pos + 8:pos + 16])[0]
            body = pos + 16
        elif size == 0:
            size = end - pos
        if size < body - pos or pos + size > end:
            return
        yield kind, body, pos + size
        pos += size


def _geojp2(data: bytes):
    """GeoJP2: a 'uuid' box whose payload is a degenerate GeoTIFF carrying the three georeferencing tags."""
    for kind, a, b in _jp2_boxes(data):
        if kind == b"uuid" and data[a:a + 16] == _GEOJP2_UUID:
            tags = _read_ifd(data[a + 16:b])
            if TAG_PIXEL_SCALE in tags and TAG_TIEPOINT in tags:
                sx, sy = float(tags[TAG_PIXEL_SCALE][0]), float(tags[TAG_PIXEL_SCALE][1])
                tp = tags[TAG_TIEPOINT]
                left, top = float(tp[3]) - float(tp[0]) * sx, float(tp[4]) + float(tp[1]) * sy
                proj, geog, point = None, None, False
                gk = tags.get(TAG_GEOKEYS, ())
                for i in range(4, len(gk) - 3, 4):
                    if gk[i + 1] != 0:
                        continue
                    if gk[i] == 3072:
                        proj = gk[i + 3]
                    elif gk[i] == 2048:
                        geog = gk[i + 3]
                    elif gk[i] == 1025:
                        point = gk[i + 3] == 2
                crs = f"EPSG:{proj or geog}" if (proj or geog) else None
                if point:                                   # PixelIsPoint: the tie point is a pixel centre
                    left, top = left - 0.5 * sx, top + 0.5 * sy
                return left, top, sx, sy, crs
    return None


def _gmljp2(data: bytes):
    """GMLJP2: an 'asoc' box tree holding an 'xml ' box with a gml:RectifiedGrid -- origin = CENTRE of the first pixel,
    two axis-aligned offset vectors (OGC 05-047r3)."""
    def xml_boxes(a, b):
        for kind, x, y in _jp2_boxes(data, a, b):
            if kind == b"asoc":
                yield from xml_boxes(x, y)
            elif kind == b"xml ":
                yield data[x:y]
    for raw in xml_boxes(0, len(data)):
        text = raw.decode("utf-8", errors="replace")
        if "RectifiedGrid" not in text:
            continue
        num = r"[-+0-9.eE]+"
        m_o = re.search(rf"<gml:origin>.*?<gml:(?:pos|coordinates)[^>]*>\s*({num})[\s,]+({num})", text, flags=re.S)
        vecs = re.findall(rf"<gml:offsetVector[^>]*>\s*({num})[\s,]+({num})", text)
        if not m_o or len(vecs) < 2:
            continue
        ox, oy = float(m_o.group(1)), float(m_o.group(2))
        (ax, ay), (bx, by) = [(float(u), float(v)) for u, v in vecs[:2]]
        if ay != 0.0 or bx != 0.0:
            if ax == 0.0 and by == 0.0:                     # axis order northing / easting (e.g. urn:...:EPSG::4326)
                ox, oy, ax, by = oy, ox, bx, ay
            else:
                raise ValueError("GMLJP2: rotated grids are not supported")
        sx, sy = abs(ax), abs(by)
        m_c = re.search(r"srsName=\"[^\"]*?EPSG[:/]+(?:[0-9.]*[:/])?(\d+)\"", text)
        return ox - 0.5 * sx, oy + 0.5 * sy, sx, sy, (f"EPSG:{m_c.group(1)}" if m_c else None)
    return None


def _world_file(path: str):
    """ESRI world file next to the image (.j2w / .jpw / .wld ...): pixel size x, rotations, -pixel size y, then the map
    coordinates of the CENTRE of the upper-left pixel."""
    base, ext = os.path.splitext(path)
    e = ext.lstrip(".")
    for cand in ((base + "." + e[0] + e[-1] + "w") if len(e) >= 2 else None, base + ".j2w", base + ".wld", path + "w"):
        if cand and os.path.isfile(cand):
            vals = [float(x) for x in open(cand).read().split()[:6]]
            if len(vals) == 6:
                a, d, b, e_, c, f = vals
                if d != 0.0 or b != 0.0:
                    raise ValueError(f"{cand}: rotated rasters are not supported")
                return c - 0.5 * a, f + 0.5 * abs(e_), abs(a), abs(e_), None
    return None


def _tab_file(path: str):
    """MapInfo .tab raster registration next to the image (how IGN delivers BD ORTHO tiles): control points
    ``(x,y) (col,row) Label "Pt n"`` -- map coordinates of pixel CORNERS -- and optionally ``CoordSys Earth Projection ...``.
    The points must describe a north-up, unrotated grid; the Lambert-93 definition is recognised, other projections leave
    the CRS unknown."""
    base = os.path.splitext(path)[0]
    for cand in (base + ".tab", base + ".TAB"):
        if not os.path.isfile(cand):
            continue
        text = open(cand, encoding="latin-1").read()
        num = r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?"
        pts = re.findall(rf"\(\s*({num})\s*,\s*({num})\s*\)\s*\(\s*({num})\s*,\s*({num})\s*\)\s*Label", text)
        if len(pts) < 2:
            continue
        p = np.asarray(pts, dtype=np.float64)                 # x, y, col, row
        cols, rows = p[:, 2], p[:, 3]
        if np.ptp(cols) == 0 or np.ptp(rows) == 0:
            continue
        i0, i1 = int(np.argmin(cols)), int(np.argmax(cols))
        j0, j1 = int(np.argmin(rows)), int(np.argmax(rows))
        sx = (p[i1, 0] - p[i0, 0]) / (cols[i1] - cols[i0])
        sy = (p[j0, 1] - p[j1, 1]) / (rows[j1] - rows[j0])
        left, top = p[i0, 0] - cols[i0] * sx, p[j0, 1] + rows[j0] * sy
        fit_x, fit_y = left + cols * sx, top - rows * sy
        if sx <= 0 or sy <= 0 or np.abs(fit_x - p[:, 0]).max() > 1e-3 * sx or np.abs(fit_y - p[:, 1]).max() > 1e-3 * sy:
            raise ValueError(f"{cand}: the control points do not describe a north-up, unrotated grid")
        crs = None
        m = re.search(r"CoordSys\s+Earth\s+Projection\s+([^\n]*)", text)
        if m:
            f = [t.strip().strip('"') for t in m.group(1).split(",")]
            try:                                              # Lambert Conformal Conic, RGF93, 46.5 / 44 / 49, 700000 / 6600000
                if int(f[0]) == 3 and int(f[1]) == 33 and [float(v) for v in f[3:9]] == [3.0, 46.5, 44.0, 49.0, 700000.0, 6600000.0]:
                    crs = "EPSG:2154"
            except (ValueError, IndexError):
                pass
        return float(left), float(top), float(sx), float(sy), crs
    return None


def _jp2_georef(path: str):
    with open(path, "rb") as f:
        head = f.read(12)
        if head[:4] == b"\xff\x4f\xff\x51":                  # a bare codestream (.j2k): no boxes, side files only
            geo = _world_file(path) or _tab_file(path)
            if geo is None:
                raise ValueError(f"{path}: a JPEG 2000 codestream carries no georeferencing; no world file or .tab beside it")
            return geo[0], geo[1], _square_res(path, geo[2], geo[3]), geo[4]
        if head[4:8] != b"jP  ":
            raise ValueError(f"{path}: not a JP2 file (no signature box)")
        # georeferencing boxes sit in front of the codestream: read up to it, not the pixels
        f.seek(0)
        meta = bytearray()
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                break
            size, kind = struct.unpack(">I4s", hdr)
            ext = b""
            if size == 1:
                ext = f.read(8)
                size = struct.unpack(">Q", ext)[0]
            if kind == b"jp2c" or size == 0:
                break
            body = f.read(size - 8 - len(ext))
            meta += hdr + ext + body
    geo = _geojp2(bytes(meta)) or _gmljp2(bytes(meta)) or _world_file(path) or _tab_file(path)
    if geo is None:
        raise ValueError(f"{path}: no georeferencing (GeoJP2 box, GMLJP2 box, world file or MapInfo .tab)")
    left, top, sx, sy, crs = geo
    return left, top, _square_res(path, sx, sy), crs


def _jp2_open(path: str):
    from PIL import Image, features
    if not features.check("jpg_2000"):
        raise ValueError(f"{path}: this Pillow build has no JPEG 2000 decoder (OpenJPEG)")
    Image.MAX_IMAGE_PIXELS = None
    return Image.open(path)


def jp2_header(path: str):
    """-> (shape (count, H, W), dtype, left, top, res, crs): boxes and the codestream's SIZ marker only."""
    left, top, res, crs = _jp2_georef(path)
    with _jp2_open(path) as im:
        if im.mode not in ("L", "RGB", "RGBA", "RGBX", "LA"):
            raise ValueError(f"{path}: only 8-bit JPEG 2000 rasters are supported (Pillow mode {im.mode})")
        shape = (len(im.getbands()), im.size[1], im.size[0])
    return shape, np.dtype(np.uint8), left, top, res, crs


def read_jp2(path: str, alloc: Optional[Allocator] = None) -> Tuple[np.ndarray, float, float, float, Optional[str]]:
    """JPEG-2000 raster (the reference's product inputs, inference.py:60 ``*.jp2``) -> (array (count, H, W) uint8, left, top,
    res, crs).  Pixels: the OpenJPEG library driven directly (``..openjpeg``: its worker threads on all cores, decoding into
    the upload buffer; row ranges for strips and progressive loads), Pillow's one-thread path as the fallback -- the image is
    decoded once, it becomes HBM-resident anyway, instead of one windowed decode per tile (dataset.py:108-115).  Georeferencing, in GDAL's order of preference: GeoJP2 uuid box, GMLJP2,
    world file."""
    left, top, res, crs = _jp2_georef(path)
    from .. import openjpeg
    try:                                                     # OpenJPEG driven directly: all cores, straight into the buffer
        if os.environ.get("FZ_JP2_DECODER", "").lower() == "pillow":
            raise openjpeg.OpenJPEGUnavailable("FZ_JP2_DECODER=pillow")
        j = openjpeg.info(path)
        shape = (j.count, j.height, j.width)
        out = alloc(shape, np.dtype(np.uint8)) if alloc is not None else np.empty(shape, np.uint8)
        openjpeg.read_rows(path, 0, j.height, out=out)
        return out, left, top, res, crs
    except Exception:  # noqa: BLE001
        pass                                                 # Pillow's one-thread path decodes what is left (or says why not)
    with _jp2_open(path) as im:
        a = np.asarray(im)
    if a.dtype != np.uint8:
        raise ValueError(f"{path}: only 8-bit JPEG 2000 rasters are supported (got {a.dtype})")
    a = a[None] if a.ndim == 2 else a.transpose(2, 0, 1)
    out = alloc(a.shape, a.dtype) if alloc is not None else np.empty(a.shape, a.dtype)
    np.copyto(out, a)                                        # (H,W,C) -> (C,H,W), into the upload buffer
    return out, left, top, res, crs


_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 12: "d", 16: "Q"}


def _read_ifd(data: bytes) -> dict:
    """Tags of the first directory of a (small, classic) TIFF held in memory: the GeoJP2 payload."""
    if data[:2] not in (b"II", b"MM"):
        raise ValueError("GeoJP2 box: not a TIFF")
    e = "<" if data[:2] == b"II" else ">"
    if struct.unpack(e + "H", data[2:4])[0] != 42:
        raise ValueError("GeoJP2 box: not a classic TIFF")
    off = struct.unpack(e + "I", data[4:8])[0]
    n = struct.unpack(e + "H", data[off:off + 2])[0]
    raw = data[off + 2:off + 2 + 12 * n]
    tags = {}
    for i in range(n):
        tag, typ, cnt = struct.unpack(e + "HHI", raw[12 * i:12 * i + 8])
        fmt = _TYPE_FMT.get(typ)
        if fmt is None or typ in (2, 5):
            continue
        size = struct.calcsize(fmt) * cnt
        if size <= 4:
            val = raw[12 * i + 8:12 * i + 8 + size]
        else:
            at = struct.unpack(e + "I", raw[12 * i + 8:12 * i + 12])[0]
            val = data[at:at + size]
        tags[tag] = struct.unpack(e + f"{cnt}{fmt}", val)
    return tags

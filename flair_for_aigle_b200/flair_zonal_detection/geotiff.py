"""Minimal GeoTIFF I/O for the two ends of the zonal path, on the host (numpy + zlib + Pillow/libtiff; rasterio/GDAL
are not required).  Replaces, for north-up rasters in a projected CRS:

  rasterio.open(path).read() / .bounds / .res / .crs      flair_zonal_detection/dataset.py:89-117, inference.py:76-132
  rasterio.open(path, 'w', **profile).write(...)          flair_zonal_detection/inference.py:157-208,343-352

Georeferencing is carried by ModelPixelScaleTag (33550), ModelTiepointTag (33922) and a GeoKeyDirectoryTag (34735)
holding the EPSG code, which is what GDAL writes for such rasters.  Single-band outputs are LZW-compressed through
Pillow's libtiff binding (the reference's ``compress='lzw'``); multi-band outputs (``class_prob``: one band per class)
are written by a small strip writer with Deflate compression and PlanarConfiguration = separate, because Pillow has no
N-band uint8 mode.  JPEG-2000 inputs need a real decoder and stay out of scope.
"""
from __future__ import annotations

import re
import struct
import zlib
from typing import Optional, Tuple

import numpy as np

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
    """arr uint8 (count, H, W).  Returns the path written."""
    assert arr.ndim == 3 and arr.dtype == np.uint8, (arr.shape, arr.dtype)
    count, h, w = arr.shape
    scale = (float(res), float(res), 0.0)
    tie = (0.0, 0.0, 0.0, float(left), float(top), 0.0)
    keys = _geokeys(crs)
    if count == 1:
        from PIL import Image, TiffImagePlugin
        Image.MAX_IMAGE_PIXELS = None
        ifd = TiffImagePlugin.ImageFileDirectory_v2()
        ifd[TAG_PIXEL_SCALE] = scale
        ifd.tagtype[TAG_PIXEL_SCALE] = 12           # DOUBLE
        ifd[TAG_TIEPOINT] = tie
        ifd.tagtype[TAG_TIEPOINT] = 12
        ifd[TAG_GEOKEYS] = keys
        ifd.tagtype[TAG_GEOKEYS] = 3                # SHORT
        Image.fromarray(arr[0]).save(path, format="TIFF", compression="tiff_lzw", tiffinfo=ifd)
        return path
    _write_planar_deflate(path, arr, scale, tie, keys)
    return path


def _write_planar_deflate(path, arr, scale, tie, keys, rows_per_strip: int = 256) -> None:
    count, h, w = arr.shape
    strips = []
    for b in range(count):
        for r0 in range(0, h, rows_per_strip):
            strips.append(zlib.compress(np.ascontiguousarray(arr[b, r0:r0 + rows_per_strip]).tobytes(), 1))
    n_strips = len(strips)
    total = 8 + sum(len(s) for s in strips)
    if total + 64 * 1024 + 8 * n_strips >= 2 ** 32:
        raise ValueError("raster too large for classic TIFF (BigTIFF is not implemented)")
    offsets, pos = [], 8
    for s in strips:
        offsets.append(pos)
        pos += len(s)
    # out-of-line values, then the IFD
    extra = bytearray()
    extra_base = pos

    def put(data: bytes) -> int:
        off = extra_base + len(extra)
        extra.extend(data)
        if len(extra) % 2:
            extra.append(0)
        return off

    entries = []

    def ent(tag, typ, n, value_bytes):
        if len(value_bytes) <= 4:
            entries.append((tag, typ, n, value_bytes.ljust(4, b"\0")))
        else:
            entries.append((tag, typ, n, struct.pack("<I", put(value_bytes))))

    ent(256, 4, 1, struct.pack("<I", w))
    ent(257, 4, 1, struct.pack("<I", h))
    ent(258, 3, count, struct.pack(f"<{count}H", *([8] * count)))
    ent(259, 3, 1, struct.pack("<H", 8))                       # Adobe Deflate
    ent(262, 3, 1, struct.pack("<H", 1))                       # BlackIsZero
    ent(273, 4, n_strips, struct.pack(f"<{n_strips}I", *offsets))
    ent(277, 3, 1, struct.pack("<H", count))
    ent(278, 4, 1, struct.pack("<I", rows_per_strip))
    ent(279, 4, n_strips, struct.pack(f"<{n_strips}I", *[len(s) for s in strips]))
    ent(284, 3, 1, struct.pack("<H", 2))                       # PlanarConfiguration = separate
    if count > 1:
        ent(338, 3, count - 1, struct.pack(f"<{count - 1}H", *([0] * (count - 1))))   # ExtraSamples: unspecified
    ent(339, 3, count, struct.pack(f"<{count}H", *([1] * count)))                     # SampleFormat: unsigned
    ent(TAG_PIXEL_SCALE, 12, 3, struct.pack("<3d", *scale))
    ent(TAG_TIEPOINT, 12, 6, struct.pack("<6d", *tie))
    ent(TAG_GEOKEYS, 3, len(keys), struct.pack(f"<{len(keys)}H", *keys))
    entries.sort(key=lambda e: e[0])
    ifd_off = extra_base + len(extra)
    ifd = struct.pack("<H", len(entries)) + b"".join(struct.pack("<HHI", t, ty, n) + v for t, ty, n, v in entries)
    ifd += struct.pack("<I", 0)
    with open(path, "wb") as f:
        f.write(b"II*\0" + struct.pack("<I", ifd_off))
        for s in strips:
            f.write(s)
        f.write(bytes(extra))
        f.write(ifd)


def read_geotiff(path: str) -> Tuple[np.ndarray, float, float, float, Optional[str]]:
    """-> (array (count, H, W), left, top, res, crs).  Pillow/libtiff decodes 1-, 3- and 4-band 8-bit images (strips or
    tiles, any libtiff codec); planar multi-band files written by ``write_geotiff`` are decoded here."""
    tags = _read_ifd(path)
    if TAG_PIXEL_SCALE not in tags or TAG_TIEPOINT not in tags:
        raise ValueError(f"{path}: no GeoTIFF georeferencing (ModelPixelScale / ModelTiepoint tags)")
    sx, sy = float(tags[TAG_PIXEL_SCALE][0]), float(tags[TAG_PIXEL_SCALE][1])
    if abs(sx - sy) > 1e-9 * max(sx, sy):
        raise ValueError(f"{path}: non-square pixels ({sx} x {sy}) are not supported")
    tp = tags[TAG_TIEPOINT]
    left, top = float(tp[3]) - float(tp[0]) * sx, float(tp[4]) + float(tp[1]) * sy
    crs = None
    gk = tags.get(TAG_GEOKEYS)
    if gk is not None:
        for i in range(4, len(gk) - 3, 4):
            if gk[i] in (3072, 2048) and gk[i + 1] == 0:
                crs = f"EPSG:{gk[i + 3]}"
    planar, comp = tags.get(284, (1,))[0], tags.get(259, (1,))[0]
    count = tags.get(277, (1,))[0]
    if planar == 2 and count > 1 and comp == 8:
        arr = _read_planar_deflate(path, tags)
    else:
        from PIL import Image
        Image.MAX_IMAGE_PIXELS = None
        with Image.open(path) as im:
            a = np.asarray(im)
        if a.dtype != np.uint8:
            raise ValueError(f"{path}: only 8-bit rasters are supported (got {a.dtype})")
        arr = a[None] if a.ndim == 2 else np.ascontiguousarray(a.transpose(2, 0, 1))
    return arr, left, top, sx, crs


_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 12: "d", 16: "Q"}


def _read_ifd(path: str) -> dict:
    with open(path, "rb") as f:
        head = f.read(8)
        if head[:2] not in (b"II", b"MM"):
            raise ValueError(f"{path}: not a TIFF file")
        e = "<" if head[:2] == b"II" else ">"
        if struct.unpack(e + "H", head[2:4])[0] != 42:
            raise ValueError(f"{path}: BigTIFF is not supported")
        f.seek(struct.unpack(e + "I", head[4:8])[0])
        n = struct.unpack(e + "H", f.read(2))[0]
        raw = f.read(12 * n)
        tags = {}
        for i in range(n):
            tag, typ, cnt = struct.unpack(e + "HHI", raw[12 * i:12 * i + 8])
            fmt = _TYPE_FMT.get(typ)
            if fmt is None or typ in (2, 5):
                continue
            size = struct.calcsize(fmt) * cnt
            if size <= 4:
                data = raw[12 * i + 8:12 * i + 8 + size]
            else:
                here = f.tell()
                f.seek(struct.unpack(e + "I", raw[12 * i + 8:12 * i + 12])[0])
                data = f.read(size)
                f.seek(here)
            tags[tag] = struct.unpack(e + f"{cnt}{fmt}", data)
    return tags


def _read_planar_deflate(path: str, tags: dict) -> np.ndarray:
    w, h, count = tags[256][0], tags[257][0], tags[277][0]
    rps = tags[278][0]
    per_band = (h + rps - 1) // rps
    out = np.empty((count, h, w), np.uint8)
    with open(path, "rb") as f:
        for i, (off, n) in enumerate(zip(tags[273], tags[279])):
            b, r0 = divmod(i, per_band)
            r0 *= rps
            f.seek(off)
            rows = min(rps, h - r0)
            out[b, r0:r0 + rows] = np.frombuffer(zlib.decompress(f.read(n)), np.uint8).reshape(rows, w)
    return out

"""Oracle (test infrastructure): rasterio's windowed, boundless, RESAMPLED read.

The reference reads every modality with
    reader.read(indexes, window=from_bounds(*tile_bounds, transform), out_shape=(C, ps, ps),
                resampling=Resampling.bilinear, boundless=True, fill_value=0)      (dataset.py:97-115)
where ``ps = patch_sizes[mod] = int(round(P / (mod_res / ref_res)))`` (model_utils.py:19-35).  For the
reference modality the window is ps x ps pixels on integer offsets (up to 1e-7 px of affine rounding)
and the read is a plain copy.  For a modality with another pixel size (DEM at 1 m under a 0.2 m
ortho: window 102.4 px -> ps 102) the window is fractional and GDAL resamples.

rasterio/GDAL are absent from this image and from /root/reference, so this restates the published
behaviour of rasterio 1.4.3 + GDAL's ``GDALRasterBand::RasterIOResampled`` / ``GDALResampleChunk``
convolution for ``bilinear``:
  * boundless: the window is served from a VRT whose background is ``fill_value``; fill pixels take
    part in the interpolation like data (no nodata is declared);
  * destination pixel i has its centre at source coordinate  off + (i + 0.5) * (win / out);
  * the kernel is the triangle  w(t) = max(0, 1 - |t|)  on source-pixel-centre distances; when the
    read downsamples (win / out > 1) the kernel is widened by that ratio (GDAL's anti-aliasing
    convolution), weights are normalised to sum 1;
  * integer output types are rounded half up (+0.5, truncation) and clamped; float types are not.
PARITY UNPINNED for the fractional case (no GDAL here to confirm the last ulp); the integer-aligned
case is exact by construction and is what every BASELINE.json config exercises.
"""
from __future__ import annotations

import math

import numpy as np


def axis_weights(off: float, win: float, n_out: int, n_src: int, method: str = "bilinear"):
    """Per output index: (first source index, weights[k]) as dense arrays.
    Returns (idx0 int64 (n_out,), w float64 (n_out, taps)).  Source indices outside [0, n_src) mean
    'fill value'."""
    ratio = win / n_out
    if method == "nearest":
        idx = np.floor(off + (np.arange(n_out) + 0.5) * ratio).astype(np.int64)
        return idx, np.ones((n_out, 1))
    support = max(ratio, 1.0)
    taps = int(math.ceil(2 * support)) + 1
    centre = off + (np.arange(n_out) + 0.5) * ratio              # source coordinate (pixel edges at integers)
    first = np.floor(centre - 0.5 - support).astype(np.int64) + 1   # first source pixel whose centre is within support
    k = np.arange(taps)[None, :]
    src_centre = first[:, None] + k + 0.5
    w = np.maximum(0.0, 1.0 - np.abs(src_centre - centre[:, None]) / support)
    s = w.sum(axis=1, keepdims=True)
    w = np.where(s > 0, w / np.where(s > 0, s, 1.0), 0.0)
    return first, w


def read_resampled(src: np.ndarray, row_off: float, col_off: float, height: float, width: float,
                   out_h: int, out_w: int, fill_value=0, method: str = "bilinear") -> np.ndarray:
    """src (C,H,W) -> (C,out_h,out_w), same dtype."""
    c, h, w = src.shape
    aligned = (abs(row_off - round(row_off)) < 1e-6 and abs(col_off - round(col_off)) < 1e-6
               and abs(height - out_h) < 1e-6 and abs(width - out_w) < 1e-6)
    if aligned:
        r0, c0 = int(round(row_off)), int(round(col_off))
        out = np.full((c, out_h, out_w), fill_value, dtype=src.dtype)
        ra, rb = max(r0, 0), min(r0 + out_h, h)
        ca, cb = max(c0, 0), min(c0 + out_w, w)
        if rb > ra and cb > ca:
            out[:, ra - r0:rb - r0, ca - c0:cb - c0] = src[:, ra:rb, ca:cb]
        return out
    iy, wy = axis_weights(row_off, height, out_h, h, method)
    ix, wx = axis_weights(col_off, width, out_w, w, method)

    def gather(axis_len, first, taps):
        idx = first[:, None] + np.arange(taps)[None, :]
        valid = (idx >= 0) & (idx < axis_len)
        return np.clip(idx, 0, axis_len - 1), valid

    yi, yv = gather(h, iy, wy.shape[1])
    xi, xv = gather(w, ix, wx.shape[1])
    s = src.astype(np.float64)
    fill = float(fill_value)
    # rows first: (C, out_h, W)
    rows = np.zeros((c, out_h, w))
    for k in range(wy.shape[1]):
        v = np.where(yv[:, k][None, :, None], s[:, yi[:, k], :], fill)
        rows += wy[:, k][None, :, None] * v
    # a column outside the raster is fill for every row
    out = np.zeros((c, out_h, out_w))
    for k in range(wx.shape[1]):
        v = np.where(xv[:, k][None, None, :], rows[:, :, xi[:, k]], fill)
        out += wx[:, k][None, None, :] * v
    if src.dtype.kind in "ui":
        info = np.iinfo(src.dtype)
        out = np.clip(np.floor(out + 0.5), info.min, info.max)
    return out.astype(src.dtype)

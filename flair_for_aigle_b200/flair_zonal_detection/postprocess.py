"""Drop-in for flair_zonal_detection/postprocess.py: ``convert`` (:9-30), computed on the GPU, and ``convert_to_cog``
(:33-52), the raster re-encoding step, on all host cores.

``convert`` accepts what the reference accepts -- a (C,H,W) array of logits -- as numpy or torch; numpy in
gives numpy out.  There is no CPU implementation: without a CUDA device it raises.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import native as nv


def convert(img, img_type: str):
    if img_type not in ("class_prob", "argmax"):
        raise ValueError(f"Unknown output type: {img_type}")
    was_numpy = isinstance(img, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(img)) if was_numpy else img
    if img_type == "class_prob" and t.dim() != 3:
        raise ValueError("Expected logits with shape (C, H, W)")
    if not torch.cuda.is_available():
        raise nv.NativeError("convert() runs on the GPU only (no CPU fallback)")
    t = t.to(device="cuda", dtype=torch.float32).contiguous()
    out = nv.convert(t, 0 if img_type == "argmax" else 1)
    return out.cpu().numpy() if was_numpy else out


def convert_to_cog(input_path: str, output_path: str) -> None:
    """postprocess.py:33-52: GeoTIFF -> Cloud Optimized GeoTIFF with the reference's profile (LZW, blocksize 512, nearest
    overviews, tiled), then the input file is removed.  The reference hands this to GDAL's COG driver through ``rio_copy``
    (single-threaded); here ``libfz_rasterio.so`` decodes, builds the overview pyramid and encodes block-parallel
    (include/flair_zonal_rasterio.h: fzio_convert_to_cog)."""
    from .. import raster_io
    if not os.path.isfile(input_path):
        raise FileNotFoundError(f"Input file not found: {input_path}")
    raster_io.convert_to_cog(input_path, output_path)
    os.remove(input_path)


def create_polygon_from_bounds(x_min: float, x_max: float, y_min: float, y_max: float) -> dict:
    """postprocess.py:55-66: ``mapping(box(x_min, y_max, x_max, y_min))`` -- the GeoJSON-like polygon of a bounding box, with
    the vertex order shapely's ``box`` gives for those arguments (its ``miny`` is y_max here) and tuples like ``mapping``."""
    ring = ((float(x_max), float(y_max)), (float(x_max), float(y_min)), (float(x_min), float(y_min)),
            (float(x_min), float(y_max)), (float(x_max), float(y_max)))
    return {"type": "Polygon", "coordinates": (ring,)}

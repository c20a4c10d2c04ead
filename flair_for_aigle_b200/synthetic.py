"""Seeded synthetic inputs shared by bench.py and the tests: a 4-band
uint8 "aerial" raster made of low-resolution noise bilinearly upsampled (land-cover-like smooth
regions) plus per-pixel noise, generated strip by strip so 60k x 60k never needs a float copy."""
from __future__ import annotations

import numpy as np


def synthetic_raster(height: int, width: int, bands: int = 4, seed: int = 2025, cell: int = 96,
                     noise: float = 6.0) -> np.ndarray:
    """uint8 (bands, height, width).  Deterministic in (shape, seed, cell, noise)."""
    rng = np.random.default_rng(seed)
    gh, gw = height // cell + 3, width // cell + 3
    coarse = rng.uniform(20.0, 235.0, size=(bands, gh, gw)).astype(np.float32)
    out = np.empty((bands, height, width), dtype=np.uint8)
    xs = (np.arange(width, dtype=np.float32) + 0.5) / cell
    x0 = np.floor(xs).astype(np.int64)
    fx = (xs - x0).astype(np.float32)
    strip = 512
    for r0 in range(0, height, strip):
        r1 = min(r0 + strip, height)
        ys = (np.arange(r0, r1, dtype=np.float32) + 0.5) / cell
        y0 = np.floor(ys).astype(np.int64)
        fy = (ys - y0).astype(np.float32)[None, :, None]
        top = coarse[:, y0][:, :, x0] * (1 - fx) + coarse[:, y0][:, :, x0 + 1] * fx
        bot = coarse[:, y0 + 1][:, :, x0] * (1 - fx) + coarse[:, y0 + 1][:, :, x0 + 1] * fx
        val = top * (1 - fy) + bot * fy
        val += rng.standard_normal(val.shape, dtype=np.float32) * noise
        out[:, r0:r1] = np.clip(np.rint(val), 0, 255).astype(np.uint8)
    return out

"""Seeded synthetic inputs shared by bench.py and the tests: a 4-band
uint8 "aerial" raster made of low-resolution noise bilinearly upsampled (land-cover-like smooth
regions) plus per-pixel noise.  Any row range can be generated independently (each 512-row band
has its own noise stream), so a rank generates only its strip of a 60k x 60k zone."""
from __future__ import annotations

import numpy as np

_BAND = 512


def synthetic_raster(height: int, width: int, bands: int = 4, seed: int = 2025, cell: int = 96, noise: float = 6.0,
                     row0: int = 0, rows: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """uint8 (bands, rows, width): rows [row0, row0+rows) of the (height x width) zone.
    Deterministic in (height, width, bands, seed, cell, noise) and independent of the row range."""
    rows = height - row0 if rows is None else rows
    gh, gw = height // cell + 3, width // cell + 3
    coarse = np.random.default_rng(seed).uniform(20.0, 235.0, size=(bands, gh, gw)).astype(np.float32)
    if out is None:
        out = np.empty((bands, rows, width), dtype=np.uint8)
    xs = (np.arange(width, dtype=np.float32) + 0.5) / cell
    x0 = np.floor(xs).astype(np.int64)
    fx = (xs - x0).astype(np.float32)
    first_band = row0 // _BAND
    last_band = (row0 + rows - 1) // _BAND
    for band in range(first_band, last_band + 1):
        b0 = band * _BAND
        b1 = min(b0 + _BAND, height)
        ys = (np.arange(b0, b1, dtype=np.float32) + 0.5) / cell
        y0 = np.floor(ys).astype(np.int64)
        fy = (ys - y0).astype(np.float32)[None, :, None]
        cy0, cy1 = coarse[:, y0], coarse[:, y0 + 1]
        top = cy0[:, :, x0] * (1 - fx) + cy0[:, :, x0 + 1] * fx
        bot = cy1[:, :, x0] * (1 - fx) + cy1[:, :, x0 + 1] * fx
        val = top * (1 - fy) + bot * fy
        rng = np.random.default_rng([seed, band])
        val += rng.standard_normal(val.shape, dtype=np.float32) * noise
        u8 = np.clip(np.rint(val), 0, 255).astype(np.uint8)
        lo, hi = max(b0, row0), min(b1, row0 + rows)
        out[:, lo - row0:hi - row0] = u8[:, lo - b0:hi - b0]
    return out


DEFAULT_MEANS = [105.66, 111.35, 102.18, 106.59]   # configs/config_model_zonal_segmentation.yaml:48 +
DEFAULT_STDS = [52.23, 45.62, 44.30, 39.78]        # configs/train/config_modalities.yaml:55-56 (IR band)


def randomize_state_(state_dict, seed: int = 2025, bf16_exact: bool = True) -> None:
    """Fill a reference-layout state_dict with seeded random values that keep activations O(1)
    through the network and leave no path numerically inert (GRN gamma/beta, biases, BatchNorm
    statistics all non-trivial).  Values are snapped to bf16-representable numbers so the bf16
    kernels and an fp32 evaluation see identical weights."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for k in sorted(state_dict.keys()):
        t = state_dict[k]
        if not t.dtype.is_floating_point:
            continue
        shape = t.shape
        if k.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif ".grn." in k:
            v = torch.randn(shape, generator=g) * (0.5 if k.endswith("weight") else 0.1)
        elif k.endswith("relative_position_bias_table"):
            v = torch.randn(shape, generator=g) * 0.5
        elif k.endswith("bias"):
            v = torch.randn(shape, generator=g) * 0.1
        elif t.dim() == 1:                       # LayerNorm / BatchNorm scale
            v = torch.rand(shape, generator=g) + 0.5
        elif t.dim() == 2:                       # linear
            v = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
            if "mlp.fc2" in k:
                v = v * 0.5
        else:                                    # conv
            fan_in = shape[1] * shape[2] * shape[3]
            v = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        if bf16_exact:
            v = v.to(torch.bfloat16).to(torch.float32)
        t.copy_(v)


def synthetic_raster_to_pinned(height: int, width: int, out, device, bands: int = 4, seed: int = 2025, cell: int = 96,
                               noise: float = 6.0, row0: int = 0, chunk_rows: int = 1024) -> None:
    """Same kind of content as ``synthetic_raster`` (smooth low-resolution field + per-pixel noise), generated on the
    GPU in row chunks and downloaded into ``out`` = a PINNED HOST uint8 tensor (bands, rows, width) holding rows
    [row0, row0+rows) of the zone.  The numpy generator makes ~10 MB/s; a 60 000 x 60 000 zone is 14.4 GB.
    Deterministic in (height, width, bands, seed, cell, noise) and independent of the row range and of the chunking
    (each chunk-aligned block of rows draws from its own seeded stream); NOT bit-identical to ``synthetic_raster``.
    Set-up helper of bench.py: the timed regions start from the pinned host tensor."""
    import torch
    rows = out.shape[1]
    gh, gw = height // cell + 3, width // cell + 3
    coarse = torch.from_numpy(np.random.default_rng(seed).uniform(20.0, 235.0, size=(bands, gh, gw)).astype(np.float32)).to(device)
    xs = (torch.arange(width, dtype=torch.float32, device=device) + 0.5) / cell
    x0 = xs.floor().long()
    fx = xs - x0
    gen = torch.Generator(device=device)
    blk0 = row0 // chunk_rows
    blk1 = (row0 + rows - 1) // chunk_rows
    for blk in range(blk0, blk1 + 1):
        b0, b1 = blk * chunk_rows, min((blk + 1) * chunk_rows, height)
        ys = (torch.arange(b0, b1, dtype=torch.float32, device=device) + 0.5) / cell
        y0 = ys.floor().long()
        fy = (ys - y0)[None, :, None]
        rows_lo = coarse[:, y0] * (1 - fy) + coarse[:, y0 + 1] * fy              # (bands, n, gw)
        val = rows_lo[:, :, x0] * (1 - fx) + rows_lo[:, :, x0 + 1] * fx          # (bands, n, width)
        gen.manual_seed(seed * 1000003 + blk)
        val += torch.randn(val.shape, generator=gen, device=device, dtype=torch.float32) * noise
        u8 = val.round_().clamp_(0, 255).to(torch.uint8)
        lo, hi = max(b0, row0), min(b1, row0 + rows)
        out[:, lo - row0:hi - row0].copy_(u8[:, lo - b0:hi - b0], non_blocking=False)

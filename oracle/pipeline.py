"""Oracle (test infrastructure): the reference's zonal pipeline on an in-memory raster, CPU.

Restates the data path of
  dataset.py:174-209 + :89-124   (windowed boundless read, fill 0, float64 (x-mean)/std -> fp32,
                                  the all-zeros label tensor that only carries ``img_size``)
  inference.py:278-352           (batch loop, model(inputs), logits -> numpy, per-tile margin
                                  crop, convert(argmax|class_prob), windowed write, last writer wins)
driving ``oracle.models.FlairHubOracle`` (torch fp32 eager).  This is also what bench.py times as
the "reference CPU torch path" (``cpu_baseline`` / ``--impl reference``), on the host cores.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .convert import write_tiles, write_tiles_rescaled
from .grid import Georef, generate_patches, read_tile, tile_plan


def normalize(patch: np.ndarray, means: Sequence[float], stds: Sequence[float]) -> np.ndarray:
    """norm.py:37-44 ('custom'): float64, per channel, in place."""
    x = patch.astype(np.float64)
    for i in range(x.shape[0]):
        x[i] -= means[i]
        x[i] /= stds[i]
    return x


def load_batch(raster: np.ndarray, plan: np.ndarray, idx: Sequence[int], patch: int, means, stds, task: str,
               n_cls: int, mod: str = "AERIAL_RGBI") -> Dict[str, torch.Tensor]:
    """dataset.py:174-209 + default collate."""
    tiles = [normalize(read_tile(raster, int(plan[i, 0]), int(plan[i, 1]), patch), means, stds) for i in idx]
    x = torch.tensor(np.ascontiguousarray(np.stack(tiles)), dtype=torch.float32)
    return {mod: x, task: torch.zeros((len(idx), n_cls, patch, patch), dtype=torch.float32)}


@torch.no_grad()
def run_zone(model, raster: np.ndarray, geo: Georef, patch: int, margin: int, means, stds, task: str, n_cls: int,
             batch_size: int = 8, output_type: str = "argmax", tile_indices: Optional[Sequence[int]] = None,
             device: str = "cpu", out_res: Optional[float] = None):
    """Full zone (or the listed tiles only).  Returns (out_raster uint8, seconds spent in the
    batch loop, n_tiles processed).  ``out_res`` = output_px_meters (inference.py:165-194,299-312)."""
    tiles = generate_patches(patch, margin, geo.res, geo)
    rescale = out_res is not None and abs(out_res - geo.res) > 1e-6
    plan = tile_plan(tiles, geo, patch, margin, out_res if rescale else None)
    order = list(range(len(tiles))) if tile_indices is None else list(tile_indices)
    h, w = geo.height, geo.width
    if rescale:
        left, bottom, right, top = geo.bounds
        h, w = int(round((top - bottom) / out_res)), int(round((right - left) / out_res))
    out = np.zeros((h, w), np.uint8) if output_type == "argmax" else np.zeros((n_cls, h, w), np.uint8)
    t0 = time.perf_counter()
    for s in range(0, len(order), batch_size):
        idx = order[s:s + batch_size]
        batch = load_batch(raster, plan, idx, patch, means, stds, task, n_cls)
        if device != "cpu":
            batch = {k: v.to(device) for k, v in batch.items()}
        logits, _ = model(batch)
        logits = logits[task].cpu().numpy()
        if rescale:
            write_tiles_rescaled(logits, plan[idx], margin, out, output_type, geo.res / out_res)
        else:
            write_tiles(logits, plan[idx], margin, out, output_type)
    return out, time.perf_counter() - t0, len(order)

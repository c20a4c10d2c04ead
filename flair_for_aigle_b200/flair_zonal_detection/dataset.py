"""Tile dataset: drop-in for flair_zonal_detection/dataset.py (``MultiModalSlicedDataset``).

The reference's ``__getitem__`` (dataset.py:174-209) decodes a window per tile in a worker
process and ships 28.3 MB of fp32 per tile to the GPU.  Here the dataset keeps the zone raster
resident in HBM (uint8, uploaded once from pinned memory) and an item is only the tile's index
and integer read-window origin: pixels are gathered, zero-filled and normalised on the device
(csrc/feeder.cu + the stem kernel).  A torch ``DataLoader`` over it still works (it batches the
tiny index records), and ``inference_and_write`` recognises the dataset and drives the fused
device path itself.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Optional

import numpy as np
import torch
from torch.utils.data import Dataset

from .raster import ZoneRaster, open_raster
from .slicing import ownership_windows, tile_plan

logger = logging.getLogger(__name__)


def normalization_affine(norm_cfg: Optional[Dict[str, Any]], channels: int, dtype=np.uint8):
    """(means, stds) such that ``(x - mean_c) / std_c`` is what the reference's ``_normalize_patch`` (dataset.py:119-124)
    -> ``norm`` (flair_hub/data/utils_data/norm.py:8-52) does to a patch of ``dtype``:
      no / empty ``normalization`` block -> identity;  ``custom`` -> the configured per-channel means / stds;
      ``scaling`` -> ``skimage.img_as_float`` (x / dtype max for unsigned integers, identity for floats);
      ``without`` -> identity.  An unknown type, or means / stds of different lengths, end the program like norm.py:33-40
    (``sys.exit(1)``).  One helper for every consumer (generic batches, fused stem, float-tile engines)."""
    ident = [0.0] * channels, [1.0] * channels
    if not norm_cfg:
        return ident
    kind = norm_cfg.get("type")
    if kind not in ("scaling", "custom", "without"):
        logger.info("Error: Normalization argument should be 'scaling', 'custom', or 'without'.")
        raise SystemExit(1)
    if kind == "custom":
        means, stds = list(norm_cfg.get("means") or []), list(norm_cfg.get("stds") or [])
        if len(means) != len(stds):
            logger.info("Error: If using 'custom', the provided means and stds must have the same length.")
            raise SystemExit(1)
        if len(means) < channels:
            # norm.py:42-44 indexes means[i] for every channel: an IndexError caught by its blanket except -> exit
            logger.info("Unexpected error during normalization: list index out of range")
            raise SystemExit(1)
        return [float(v) for v in means[:channels]], [float(v) for v in stds[:channels]]
    if kind == "scaling":
        dt = np.dtype(dtype)
        if dt.kind == "u":
            return [0.0] * channels, [float(np.iinfo(dt).max)] * channels
        if dt.kind == "f":
            return ident
        raise NotImplementedError(f"'scaling' normalisation of {dt} rasters")
    return ident


class MultiModalSlicedDataset(Dataset):
    def __init__(self, dataframe, modality_cfgs: Dict[str, Dict[str, Any]], patch_size_dict: Dict[str, int],
                 ref_date_str: Optional[str], modalities_config: Dict[str, Any]) -> None:
        self.df = dataframe
        self.modalities = modality_cfgs
        self.modalities_config = modalities_config
        self.patch_sizes = patch_size_dict
        self.ref_date_str = ref_date_str
        if any(m.endswith("_TS") for m in modality_cfgs):
            raise NotImplementedError("Sentinel time-series modalities are outside the zonal hot path")
        self.readers: Dict[str, ZoneRaster] = {m: open_raster(c['input_img_path']) for m, c in modality_cfgs.items()}
        self._device_rasters: Dict[str, torch.Tensor] = {}
        self._rows_ready: Dict[str, Any] = {}
        self._plan = None

    # ---- integer plan shared by the feeder and the writer kernels
    def plan(self) -> np.ndarray:
        if self._plan is None:
            cfg = self.modalities_config
            ref_mod = cfg.get('reference_modality', next(iter(self.readers)))
            b = self.readers[ref_mod].bounds
            ib = {'left': b.left, 'bottom': b.bottom, 'right': b.right, 'top': b.top}
            ref_res = cfg['reference_resolution']
            self._plan = tile_plan(self.df, ib, ref_res, int(cfg['img_pixels_detection']), int(cfg['margin']),
                                   cfg.get('output_px_meters', ref_res))
        return self._plan

    def host_raster(self, mod: str) -> torch.Tensor:
        """uint8 (C,H,W) in page-locked host memory: the channels listed in the modality config."""
        key = f"{mod}@host"
        if key not in self._device_rasters:
            r = self.readers[mod]
            chans = list(self.modalities[mod].get('channels') or range(1, r.count + 1))
            begin = getattr(r, 'begin_progressive', None)
            if begin is not None and chans == list(range(1, r.count + 1)) and r.profile['dtype'] == 'uint8':
                prog = begin()                           # a file: decode it in the background, bottom rows first
                if prog is not None and prog.tensor is not None:
                    self._rows_ready[mod] = prog.wait_rows
                    self._device_rasters[key] = prog.tensor
                    return prog.tensor
            arr = r.read() if chans == list(range(1, r.count + 1)) else r.read(chans)
            if arr.dtype != np.uint8:
                raise NotImplementedError(f"{mod}: only uint8 rasters are supported by the device feeder")
            pinned = getattr(r, 'pinned_tensor', None)
            if pinned is not None and arr is r.array:
                host = pinned
            else:
                host = torch.from_numpy(np.ascontiguousarray(arr))
                try:
                    host = host.pin_memory()
                except RuntimeError:  # pragma: no cover - pinning can fail on exotic hosts
                    pass
            self._device_rasters[key] = host
        return self._device_rasters[key]

    def host_rows_ready(self, mod: str):
        """``wait(lo, hi)`` that returns once rows >= lo of ``host_raster(mod)`` hold decoded pixels, or None when they all
        do already (the raster came from memory or was decoded in one go)."""
        return self._rows_ready.get(mod)

    def device_raster(self, mod: str, device) -> torch.Tensor:
        """uint8 (imagery) or float32 (elevation) (C,H,W) on ``device``: the channels listed in the modality config
        (1-based, like rasterio ``indexes``), uploaded once through pinned memory."""
        key = f"{mod}@{device}"
        if key not in self._device_rasters:
            r = self.readers[mod]
            chans = list(self.modalities[mod].get('channels') or range(1, r.count + 1))
            if chans == list(range(1, r.count + 1)):
                arr = r.read()                       # all bands in order: no host copy
            else:
                arr = r.read(chans)
            if arr.dtype not in (np.uint8, np.float32):
                raise NotImplementedError(f"{mod}: the device feeder reads uint8 or float32 rasters, got {arr.dtype}")
            pinned = getattr(r, 'pinned_tensor', None)
            if pinned is not None and arr is r.array:
                host = pinned                        # raster was created in pinned memory
            else:
                host = torch.from_numpy(np.ascontiguousarray(arr))
                try:
                    host = host.pin_memory()
                except RuntimeError:  # pragma: no cover - pinning can fail on exotic hosts
                    pass
            self._device_rasters[key] = host.to(device, non_blocking=True)
        return self._device_rasters[key]

    def modality_windows(self, mod: str) -> np.ndarray:
        """float64 [n,4] (row_off, col_off, height, width) of every tile's read window in ``mod``'s own pixel grid:
        ``from_bounds(*row.geometry.bounds, transform=reader.transform)`` of dataset.py:97."""
        r = self.readers[mod]
        g = np.asarray([t.bounds for t in self.df["geometry"]], dtype=np.float64).reshape(-1, 4)
        row0, col0, h, w = r.window_from_bounds(g[:, 0], g[:, 1], g[:, 2], g[:, 3])
        return np.ascontiguousarray(np.stack([row0, col0, h, w], axis=1))

    def modality_read_plan(self, mod: str):
        """How ``mod``'s tiles are read: ("aligned", int32 [n,2] origins) when every window starts on a whole pixel and is
        exactly ``patch_sizes[mod]`` pixels wide (the reference modality always; any modality at the reference
        resolution) -- a plain copy with zero fill; else ("resampled", float64 [n,4] windows): the window is a fractional
        number of the modality's pixels (a coarser DEM under a 0.2 m ortho) and rasterio resamples it bilinearly to
        ``patch_sizes[mod]`` (dataset.py:108-115) -> fz_gather_tiles_resampled."""
        cfg = self.modalities_config
        ref_mod = cfg.get('reference_modality', next(iter(self.readers)))
        if mod == ref_mod:
            return "aligned", np.ascontiguousarray(self.plan()[:, :2]).astype(np.int32)
        win = self.modality_windows(mod)
        ps = int(self.patch_sizes[mod])
        if win.shape[0] == 0:
            return "aligned", np.zeros((0, 2), np.int32)
        aligned = (np.abs(win[:, :2] - np.round(win[:, :2])).max() < 1e-6 and np.abs(win[:, 2:] - ps).max() < 1e-6)
        if aligned:
            return "aligned", np.round(win[:, :2]).astype(np.int32)
        return "resampled", win

    def modality_origins(self, mod: str) -> np.ndarray:
        """int32 [n,2] (row0, col0) of every tile's read window in ``mod``'s pixel grid, for modalities read without
        resampling (see ``modality_read_plan``)."""
        kind, plan = self.modality_read_plan(mod)
        if kind != "aligned":
            raise ValueError(f"{mod}: tile windows are fractional in this raster's pixels; use modality_read_plan()")
        return plan

    def __len__(self) -> int:
        return len(self.df)

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        p = self.plan()
        return {'index': torch.tensor([idx], dtype=torch.long),
                'origin': torch.tensor([int(p[idx, 0]), int(p[idx, 1])], dtype=torch.int32)}

"""Test infrastructure: makes the reference's OWN Python importable in this image.

``/root/reference`` is pure Python but imports rasterio / geopandas / shapely / skimage /
pytorch_lightning / torchmetrics / segmentation_models_pytorch, none of which is installed here
(no network).  ``install()`` registers minimal stand-ins for exactly the names the reference's
hot-path modules touch and puts ``/root/reference`` on ``sys.path``, so that tests can call the
reference's unmodified functions:

    flair_zonal_detection.slicing.generate_patches_from_reference      (slicing.py:20-121)
    flair_zonal_detection.postprocess.convert                          (postprocess.py:9-30)
    flair_zonal_detection.dataset.MultiModalSlicedDataset              (dataset.py:24-215)
    flair_zonal_detection.inference.{initialize_geometry_and_resolutions, prep_dataset,
        init_outputs, inference_and_write, inference, logits_to_labels_and_confidence}
    flair_zonal_detection.model_utils.{compute_patch_sizes, prepare_model_config, build_inference_model}
    flair_hub.models.flair_model.{FLAIR_HUB_Model, FusionHandler}, monotemp_model.FLAIR_Monotemp
    flair_hub.models.checkpoint.load_checkpoint
    flair_hub.data.utils_data.norm.norm
    flair_hub.tasks.tasks_module.SegmentationTask.step, module_setup.{build_segmentation_module, FLAIRLosses…}

What the stand-ins restate (third-party behaviour, absent from /root/reference; each is the
published algorithm of the pinned dependency in requirements.txt):
  * rasterio 1.4.3: ``Affine`` algebra incl. the inverse used by ``windows.from_bounds`` /
    ``features.geometry_window``; ``transform.array_bounds/from_origin``; ``mask.mask(crop=True)``
    (shape + transform of the crop only); an in-memory dataset whose ``read(window=…,
    boundless=True, fill_value=0)`` serves integer-aligned windows byte for byte and resampled
    windows through ``oracle.resample.read_resampled`` (GDAL's bilinear RasterIO restated);
    a recording writer for ``open(path, 'w')``.
  * geopandas.GeoDataFrame = pandas.DataFrame (the reference only uses ``.iloc``, column access,
    ``len``); shapely ``box`` = an object with ``.bounds``.
  * skimage.img_as_float for uint8/uint16/float inputs.
  * pytorch_lightning: ``LightningModule`` = ``nn.Module`` + ``log``/``save_hyperparameters``
    no-ops; ``rank_zero_only`` = identity.
  * segmentation_models_pytorch.create_model -> the oracle's restated smp 0.4.0 / timm modules
    (``oracle.models.make_encoder/make_decoder``): this is the one place where the reference's
    arithmetic lives in an absent dependency; everything AROUND it (model wiring, dummy-feature
    stripping, fusion, final interpolate, checkpoint loading) is then the reference's own code.

Nothing here is imported by the product package or by the ``-m gpu`` tests; ``/root/reference``
does not exist on the GPU box, where ``available()`` is False and the pin tests skip.
"""
from __future__ import annotations

import math
import os
import sys
import types
from collections import namedtuple
from typing import Dict

import numpy as np

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "flair_zonal_detection", "slicing.py"))


# ----------------------------------------------------------------------------------------------
# rasterio stand-in
# ----------------------------------------------------------------------------------------------
class Affine(tuple):
    """affine.Affine: | a b c ; d e f ; 0 0 1 |."""

    def __new__(cls, a, b, c, d, e, f):
        return tuple.__new__(cls, (float(a), float(b), float(c), float(d), float(e), float(f), 0.0, 0.0, 1.0))

    a = property(lambda s: s[0]); b = property(lambda s: s[1]); c = property(lambda s: s[2])
    d = property(lambda s: s[3]); e = property(lambda s: s[4]); f = property(lambda s: s[5])

    def __mul__(self, other):
        if isinstance(other, Affine):
            sa, sb, sc, sd, se, sf = self[:6]
            oa, ob, oc, od, oe, of = other[:6]
            return Affine(sa * oa + sb * od, sa * ob + sb * oe, sa * oc + sb * of + sc,
                          sd * oa + se * od, sd * ob + se * oe, sd * oc + se * of + sf)
        x, y = other
        return (x * self.a + y * self.b + self.c, x * self.d + y * self.e + self.f)

    def __invert__(self):
        # affine.Affine.__invert__
        det = self.a * self.e - self.b * self.d
        idet = 1.0 / det
        ra, rb, rd, re = self.e * idet, -self.b * idet, -self.d * idet, self.a * idet
        return Affine(ra, rb, -self.c * ra - self.f * rb, rd, re, -self.c * rd - self.f * re)

    @staticmethod
    def translation(x, y):
        return Affine(1.0, 0.0, x, 0.0, 1.0, y)

    @staticmethod
    def scale(x, y):
        return Affine(x, 0.0, 0.0, 0.0, y, 0.0)


def from_origin(west, north, xsize, ysize):
    # rasterio.transform.from_origin: Affine.translation(west, north) * Affine.scale(xsize, -ysize)
    return Affine.translation(west, north) * Affine.scale(xsize, -ysize)


def array_bounds(height, width, transform):
    # rasterio.transform.array_bounds (rectilinear branch)
    a, b, c, d, e, f = transform[:6]
    assert b == 0 and d == 0
    return c, f + e * height, c + a * width, f


def rowcol_float(transform, xs, ys):
    inv = ~transform
    cols, rows = [], []
    for x, y in zip(xs, ys):
        cx, ry = inv * (x, y)
        cols.append(cx)
        rows.append(ry)
    return rows, cols


class Window:
    def __init__(self, col_off, row_off, width, height):
        self.col_off, self.row_off, self.width, self.height = col_off, row_off, width, height

    def __repr__(self):
        return f"Window(col_off={self.col_off}, row_off={self.row_off}, width={self.width}, height={self.height})"


def window_from_bounds(left, bottom, right, top, transform=None, **_):
    # rasterio.windows.from_bounds (1.4): float offsets/lengths through the inverse transform
    rows, cols = rowcol_float(transform, [left, right, right, left], [top, top, bottom, bottom])
    row_start, row_stop = min(rows), max(rows)
    col_start, col_stop = min(cols), max(cols)
    return Window(col_start, row_start, max(col_stop - col_start, 0.0), max(row_stop - row_start, 0.0))


BoundingBox = namedtuple("BoundingBox", "left bottom right top")

_RASTERS: Dict[str, "MemoryDataset"] = {}
_WRITERS: Dict[str, "RecordingWriter"] = {}


class MemoryDataset:
    """What ``rasterio.open(path)`` returns for a registered in-memory raster (C,H,W)."""

    def __init__(self, array: np.ndarray, left: float, top: float, res: float, crs="EPSG:2154", name="mem"):
        assert array.ndim == 3
        self.array, self.name = array, name
        self.count, self.height, self.width = array.shape
        self.shape = (self.height, self.width)
        self.transform = from_origin(left, top, res, res)
        self.res = (self.transform.a, -self.transform.e)
        self.bounds = BoundingBox(*array_bounds(self.height, self.width, self.transform))
        self.crs = crs
        self.profile = {"driver": "GTiff", "dtype": str(array.dtype), "nodata": None, "width": self.width,
                        "height": self.height, "count": self.count, "crs": crs, "transform": self.transform}
        self.closed = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        self.closed = True

    def window_transform(self, window):
        return self.transform * Affine.translation(window.col_off, window.row_off)

    def read(self, indexes=None, window=None, out_shape=None, resampling=None, boundless=False, fill_value=0, **_):
        from oracle.resample import read_resampled
        idx = list(range(1, self.count + 1)) if indexes is None else ([indexes] if np.isscalar(indexes) else list(indexes))
        src = self.array[[i - 1 for i in idx]]
        if window is None:
            window = Window(0, 0, self.width, self.height)
        oh, ow = (out_shape[-2], out_shape[-1]) if out_shape is not None else (int(round(window.height)), int(round(window.width)))
        out = read_resampled(src, window.row_off, window.col_off, window.height, window.width, oh, ow,
                             fill_value=fill_value, method=("nearest" if resampling == Resampling.nearest else "bilinear"))
        return out[0] if np.isscalar(indexes) else out


class RecordingWriter:
    """``rasterio.open(path, 'w', **profile)``: an in-memory canvas; later writes overwrite earlier ones."""

    def __init__(self, path, **profile):
        self.path, self.profile = path, dict(profile)
        self.canvas = np.zeros((profile["count"], profile["height"], profile["width"]), dtype=profile["dtype"])
        self.writes = []
        self.closed = False

    def write(self, arr, indexes=None, window=None):
        r0, c0 = int(window.row_off), int(window.col_off)
        h, w = int(window.height), int(window.width)
        assert arr.shape == (h, w), (arr.shape, window)
        self.canvas[indexes - 1, r0:r0 + h, c0:c0 + w] = arr
        self.writes.append((indexes, r0, c0, h, w))

    def close(self):
        self.closed = True


def register_raster(path: str, ds: MemoryDataset) -> None:
    _RASTERS[path] = ds


def writer(path: str) -> RecordingWriter:
    return _WRITERS[path]


def _rio_open(path, mode="r", **profile):
    if mode == "w":
        w = RecordingWriter(path, **profile)
        _WRITERS[path] = w
        return w
    if path not in _RASTERS:
        raise FileNotFoundError(path)
    return _RASTERS[path]


class Resampling:
    nearest = 0
    bilinear = 1


def _geom_bounds(shapes):
    if hasattr(shapes, "bounds") and not isinstance(shapes, (list, tuple, np.ndarray)):
        shapes = [shapes]
    bs = [s.bounds for s in shapes]
    return bs


def rio_mask(dataset, shapes, crop=False, **_):
    """rasterio.mask.mask(..., crop=True): only the crop window matters to slicing.py:41-48.
    rasterio.features.geometry_window: pixel-space bounds of every shape through the inverse
    transform, floor the starts, ceil the stops, intersect with the raster; no overlap ->
    WindowError -> ValueError('Input shapes do not overlap raster.')."""
    assert crop
    inv = ~dataset.transform
    cols, rows = [], []
    for (minx, miny, maxx, maxy) in _geom_bounds(shapes):
        for x, y in ((minx, miny), (minx, maxy), (maxx, miny), (maxx, maxy)):
            c, r = inv * (x, y)
            cols.append(c)
            rows.append(r)
    row_start, row_stop = int(math.floor(min(rows))), int(math.ceil(max(rows)))
    col_start, col_stop = int(math.floor(min(cols))), int(math.ceil(max(cols)))
    r0, c0 = max(row_start, 0), max(col_start, 0)
    r1, c1 = min(row_stop, dataset.height), min(col_stop, dataset.width)
    if r1 <= r0 or c1 <= c0:
        raise ValueError("Input shapes do not overlap raster.")
    out_image = np.broadcast_to(np.zeros((), dataset.array.dtype), (dataset.count, r1 - r0, c1 - c0))
    return out_image, dataset.window_transform(Window(c0, r0, c1 - c0, r1 - r0))


class Box:
    """shapely.geometry.box(minx, miny, maxx, maxy): only ``.bounds`` is used (dataset.py:177)."""

    def __init__(self, minx, miny, maxx, maxy, ccw=True):
        xs, ys = (minx, maxx), (miny, maxy)
        self.bounds = (float(min(xs)), float(min(ys)), float(max(xs)), float(max(ys)))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__stub__ = True
    sys.modules[name] = m
    return m


def _have(name: str) -> bool:
    try:
        __import__(name)
        return not getattr(sys.modules[name], "__stub__", False)
    except Exception:
        return False


_INSTALLED = False


def install() -> None:
    """Idempotent.  Raises RuntimeError when /root/reference is absent (callers skip)."""
    global _INSTALLED
    if not available():
        raise RuntimeError("reference tree not present")
    if _INSTALLED:
        return
    import pandas as pd
    import torch
    from torch import nn

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.insert(0, repo)

    if not _have("rasterio"):
        rio = _module("rasterio", open=_rio_open, Affine=Affine)
        rio.mask = _module("rasterio.mask", mask=rio_mask)
        rio.transform = _module("rasterio.transform", array_bounds=array_bounds, from_origin=from_origin,
                                rowcol=lambda t, xs, ys, op=math.floor: tuple(
                                    [op(v) for v in a] for a in rowcol_float(t, np.atleast_1d(xs), np.atleast_1d(ys))),
                                Affine=Affine)
        rio.windows = _module("rasterio.windows", from_bounds=window_from_bounds, Window=Window)
        rio.enums = _module("rasterio.enums", Resampling=Resampling)
        rio.io = _module("rasterio.io", DatasetReader=MemoryDataset, DatasetWriter=RecordingWriter)
        rio.features = _module("rasterio.features", shapes=lambda *a, **k: iter(()))
        rio.shutil = _module("rasterio.shutil", copy=lambda *a, **k: None)
    if not _have("shapely"):
        sh = _module("shapely")
        sh.geometry = _module("shapely.geometry", box=Box, Polygon=Box, shape=lambda g: g, mapping=lambda g: g)
    if not _have("geopandas"):
        class GeoDataFrame(pd.DataFrame):
            def __init__(self, data=None, *args, crs=None, geometry=None, **kw):
                super().__init__(data, *args, **kw)

            @property
            def _constructor(self):
                return GeoDataFrame
        _module("geopandas", GeoDataFrame=GeoDataFrame, read_file=None, read_postgis=None)
    if not _have("skimage"):
        def img_as_float(a):
            a = np.asarray(a)
            if a.dtype.kind == "f":
                return a
            if a.dtype.kind == "u":
                return a.astype(np.float64) / np.iinfo(a.dtype).max
            if a.dtype.kind == "i":
                info = np.iinfo(a.dtype)
                return np.clip(a.astype(np.float64) / info.max, -1.0, 1.0)
            raise ValueError(a.dtype)
        _module("skimage", img_as_float=img_as_float)
    if not _have("pytorch_lightning"):
        def rank_zero_only(fn):
            return fn

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

            def log_dict(self, *a, **k):
                pass

            @property
            def device(self):
                return next(self.parameters()).device

        class LightningDataModule:
            def __init__(self, *a, **k):
                pass

        pl = _module("pytorch_lightning", LightningModule=LightningModule, LightningDataModule=LightningDataModule,
                     Trainer=object)
        pl.utilities = _module("pytorch_lightning.utilities")
        pl.utilities.rank_zero = _module("pytorch_lightning.utilities.rank_zero", rank_zero_only=rank_zero_only)
        pl.utilities.rank_zero_only = rank_zero_only
    if not _have("torchmetrics"):
        class _Metric(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

            def update(self, *a, **k):
                pass

            def forward(self, *a, **k):
                return torch.zeros(())

            def compute(self):
                return torch.zeros(())

            def reset(self):
                pass
        tm = _module("torchmetrics")
        tm.classification = _module("torchmetrics.classification", MulticlassJaccardIndex=_Metric)
        tm.aggregation = _module("torchmetrics.aggregation", MeanMetric=_Metric)
    if not _have("segmentation_models_pytorch"):
        from oracle import models as om

        NATIVE = {"resnet18", "resnet34"}     # smp 0.4.0 ``encoders`` registry entries the oracle restates

        class _SegModel(nn.Module):
            def __init__(self, encoder, wrapper):
                super().__init__()
                self.encoder, self.decoder, self.segmentation_head = encoder, wrapper.decoder, wrapper.segmentation_head

        def create_model(arch, encoder_name="resnet34", encoder_weights="imagenet", in_channels=3, classes=1, **kwargs):
            # smp.create_model -> Unet/UPerNet(encoder_name=…) -> get_encoder: names outside the native
            # registry raise KeyError unless prefixed 'tu-' (monotemp_model.py:67-92 relies on that)
            if not encoder_name.startswith("tu-") and encoder_name not in NATIVE:
                raise KeyError(f"Wrong encoder name `{encoder_name}`")
            enc = om.make_encoder(encoder_name, in_channels)
            return _SegModel(enc, om.make_decoder(arch, enc.out_channels, classes))
        _module("segmentation_models_pytorch", create_model=create_model)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    _INSTALLED = True

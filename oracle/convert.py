"""Oracle (test infrastructure): logits -> class raster post-processing.

Restates ``flair_zonal_detection/postprocess.py:9-30`` (``convert``), the per-tile
crop / write loop of ``inference.py:297-352`` (mode "write": last writer wins) and the
*intended* behaviour of the accumulating variant ``inference.py:468-572`` (mode "blend":
the reference's int8 accumulator wrap and ``top_px`` sign bug, SURVEY.md A8, are NOT
reproduced -- the accumulator is wide and the window is the one of inference.py:318-343).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
from scipy.special import softmax


def convert(img: np.ndarray, img_type: str) -> np.ndarray:
    """postprocess.py:9-30."""
    if img_type == "class_prob":
        if img.ndim != 3:
            raise ValueError("Expected logits with shape (C, H, W)")
        img = softmax(img, axis=0)
        return np.round(img * 255).astype(np.uint8)
    elif img_type == "argmax":
        prediction = np.argmax(img, axis=0)
        return np.expand_dims(prediction.astype(np.uint8), axis=0)
    else:
        raise ValueError(f"Unknown output type: {img_type}")


def write_tiles(logits: np.ndarray, plan: np.ndarray, margin: int, out: np.ndarray,
                output_type: str = "argmax") -> None:
    """inference.py:297-352 on an in-memory raster: logits (B,C,P,P) fp32, plan rows
    [row0,col0,top_px,left_px,height_px,width_px] (oracle/grid.py tile_plan), out is
    (H,W) uint8 for argmax or (C,H,W) uint8 for class_prob.  Tiles are written in order;
    later tiles overwrite earlier ones."""
    p = logits.shape[-1]
    for i in range(logits.shape[0]):
        top_px, left_px, h, w = (int(v) for v in plan[i, 2:6])
        if h <= 0 or w <= 0:
            continue
        patch = logits[i, :, margin:p - margin, margin:p - margin]
        pred = convert(patch, output_type)[..., :h, :w]
        if output_type == "argmax":
            out[top_px:top_px + h, left_px:left_px + w] = pred[0]
        else:
            out[:, top_px:top_px + h, left_px:left_px + w] = pred


def write_tiles_rescaled(logits: np.ndarray, plan: np.ndarray, margin: int, out: np.ndarray, output_type: str,
                         scale: float) -> None:
    """inference.py:297-352 with ``output_px_meters != reference_resolution``: ``resample_prediction``
    (inference.py:212-226 = scipy.ndimage.zoom, order 0) on the labels for 'argmax', on the logits before
    ``convert`` for 'class_prob'; plan is in OUTPUT pixels (oracle/grid.py tile_plan with out_res)."""
    from scipy.ndimage import zoom
    p = logits.shape[-1]
    for i in range(logits.shape[0]):
        top_px, left_px, h, w = (int(v) for v in plan[i, 2:6])
        if h <= 0 or w <= 0:
            continue
        patch = logits[i, :, margin:p - margin, margin:p - margin]
        if output_type == "argmax":
            pred = zoom(convert(patch, "argmax"), zoom=(1, scale, scale), order=0)
            out[top_px:top_px + h, left_px:left_px + w] = pred[0, :h, :w]
        else:
            pred = convert(zoom(patch, zoom=(1, scale, scale), order=0), output_type)
            out[:, top_px:top_px + h, left_px:left_px + w] = pred[:, :h, :w]


def blend_accumulate(logits: np.ndarray, plan: np.ndarray, margin: int, canvas: np.ndarray,
                     weights: Optional[np.ndarray] = None) -> None:
    """Intended semantics of inference.py:520-562: softmax over classes of the
    margin-cropped tile, accumulated (optionally weighted per pixel by ``weights`` of shape
    (P-2m, P-2m)) into ``canvas`` (C,H,W) float32.  fp32 softmax like scipy on fp32 input."""
    p = logits.shape[-1]
    for i in range(logits.shape[0]):
        top_px, left_px, h, w = (int(v) for v in plan[i, 2:6])
        if h <= 0 or w <= 0:
            continue
        patch = logits[i, :, margin:p - margin, margin:p - margin].astype(np.float32)
        prob = softmax(patch, axis=0)[:, :h, :w]
        if weights is not None:
            prob = prob * weights[None, :h, :w]
        canvas[:, top_px:top_px + h, left_px:left_px + w] += prob


def blend_accumulate_rescaled(logits: np.ndarray, plan: np.ndarray, margin: int, canvas: np.ndarray,
                              scale: float) -> None:
    """inference.py:515-562 with ``output_px_meters != reference_resolution`` and the wide accumulator:
    ``resample_prediction`` (scipy.ndimage.zoom, order 0) on the cropped logits (:521-523), ``convert``-style softmax
    (:525), accumulation at the out_res pixel position (:530-562).  plan is in OUTPUT pixels."""
    from scipy.ndimage import zoom
    p = logits.shape[-1]
    for i in range(logits.shape[0]):
        top_px, left_px, h, w = (int(v) for v in plan[i, 2:6])
        if h <= 0 or w <= 0:
            continue
        patch = zoom(logits[i, :, margin:p - margin, margin:p - margin].astype(np.float32), zoom=(1, scale, scale), order=0)
        canvas[:, top_px:top_px + h, left_px:left_px + w] += softmax(patch, axis=0)[:, :h, :w]


def logits_to_labels_and_confidence(probs: np.ndarray):
    """inference.py:566-572."""
    labels = np.argmax(probs, axis=0).astype(np.uint8)
    confidence = np.max(probs, axis=0)
    return labels, confidence

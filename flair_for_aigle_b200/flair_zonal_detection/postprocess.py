"""Drop-in for flair_zonal_detection/postprocess.py:9-30 (``convert``), computed on the GPU.

Accepts what the reference accepts -- a (C,H,W) array of logits -- as numpy or torch; numpy in
gives numpy out.  There is no CPU implementation: without a CUDA device it raises.
"""
from __future__ import annotations

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

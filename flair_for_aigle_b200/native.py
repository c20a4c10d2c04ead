"""ctypes binding of libfz_b200.so (C ABI declared in include/flair_zonal_b200.h).

PyTorch is used only as the owner of device memory and streams: every wrapper takes torch
CUDA tensors, passes raw ``data_ptr()``s and the current stream handle.  There is NO CPU or
library fallback: if the shared library is missing or a call fails, a ``NativeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Optional

import torch

# FZ_OPERANDS=bf16 selects the A/B build whose inference kernels store bf16 operands (round-1 behaviour, see
# csrc/operand.cuh and build.py); the product library stores fp16.
# FZ_LIB_VARIANT=<name> loads _native/libfz_b200_<name>.so: a hand-built A/B variant of the library (tools/build_variant.py).
_LIB_PATH = Path(__file__).resolve().parent / "_native" / (
    f"libfz_b200_{os.environ['FZ_LIB_VARIANT']}.so" if os.environ.get("FZ_LIB_VARIANT") else
    ("libfz_b200_bf16.so" if os.environ.get("FZ_OPERANDS", "").lower() == "bf16" else "libfz_b200.so"))
_lib: Optional[ctypes.CDLL] = None

F32, BF16, F16 = 0, 1, 2
ABI_VERSION = 3
EPI_OPERANDS_F16 = 0x200
NCHW, NHWC, NHWC_UP4 = 0, 1, 2
EPI_BF16, EPI_GELU_SUMSQ, EPI_RESID_F32, EPI_F32, EPI_RELU_BF16, EPI_GELU_BF16 = 0, 1, 2, 3, 4, 5
EPI_REVERSE_TILES = 0x100
CONV_RELU_BF16, CONV_LOGITS_F32, CONV_ARGMAX_RASTER, CONV_LOGITS_F32_NCHW, CONV_ADD_RELU_BF16, CONV_BF16 = 0, 1, 2, 3, 4, 5


class NativeError(RuntimeError):
    pass


# Optional per-launch timing (bench.py's roofline leg): when PROFILE is a list, every kernel wrapper
# brackets its launch with CUDA events on the launching stream and appends
# (kernel_family, meta, start_event, end_event).  Never enabled inside CUDA-graph capture.
PROFILE = None


# FZ_NVTX=1: every kernel wrapper also opens an NVTX range named after its kernel family (+ its shape), so that
# `ncu --nvtx --nvtx-include "gemm_tcgen05/"` (or any NVTX-aware profiler) can select launches by what they compute.
NVTX = os.environ.get("FZ_NVTX", "0") == "1"


class _Timed:
    def __init__(self, family: str, **meta):
        self.family, self.meta = family, meta

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.family + ("/" + ",".join(f"{k}={v}" for k, v in self.meta.items()) if self.meta else ""))
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record(torch.cuda.current_stream())
            PROFILE.append((self.family, self.meta, self.e0, self.e1))
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


_vp, _i, _i64, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "fz_last_error": [],
    "fz_abi_version": [],
    "fz_operand_format": [],
    "fz_device_info": [_i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_sz)],
    "fz_gather_tiles_f32": [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "fz_gather_tiles_f32_from_f32": [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "fz_gather_tiles_resampled": [_vp, _i, _i, _i, _i, _vp, _i, _i, ctypes.c_double, _vp, _vp, _vp, _vp],
    "fz_gather_tiles_u8": [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp],
    "fz_crop_argmax_write": [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp],
    "fz_crop_softmax_write": [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp],
    "fz_crop_softmax_accumulate": [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "fz_crop_zoom_accumulate": [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp],
    "fz_crop_zoom_write": [_i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp],
    "fz_canvas_argmax": [_vp, _i, _i64, _vp, _vp, _vp],
    "fz_convert": [_vp, _i, _i, _i, _i, _vp, _vp],
    "fz_gemm_bf16": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "fz_gemm_set_trace": [_vp],
    "fz_gemm_bf16_splitk": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_gemm_splitk_max_splits": [_i, _i, _i],
    "fz_gemm_bf16_splitk_tn": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_gemm_bf16_simt": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "fz_stem_ln": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, ctypes.c_float, _vp],
    "fz_stem_ln_f32": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, ctypes.c_float, _vp],
    "fz_dwconv7_ln": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp],
    "fz_ln2d_s2d": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp],
    "fz_ln2d_s2d_copy": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp],
    "fz_grn_scale": [_vp, _i, _vp, _vp, _vp, _i, _i, ctypes.c_float, _vp],
    "fz_scale_weights": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "fz_scale_rows": [_vp, _vp, _i64, _i, _i, _vp],
    "fz_upsample2_concat": [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "fz_catconv3x3_bn_relu": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "fz_upconv3x3_bn_relu": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "fz_conv7x7s2_bn_relu": [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "fz_maxpool3x3s2": [_vp, _vp, _i, _i, _i, _i, _vp],
    "fz_conv3x3_ex": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp],
    "fz_conv3x3_bf16": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp],
    "fz_layernorm_rows": [_vp, _vp, _vp, _vp, _i64, _i, ctypes.c_float, _vp],
    "fz_merge_ln": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, ctypes.c_float, _vp],
    "fz_swin_window_attn": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, ctypes.c_float, _vp],
    "fz_cast_f32_bf16": [_vp, _vp, _i64, _vp],
    "fz_cast_f32_16": [_vp, _vp, _i, _i64, _vp],
    "fz_adaptive_avgpool": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_bilinear_slice": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "fz_updown_slice": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "fz_pyramid_concat": [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _vp],
    "fz_ccl_label": [_vp, _vp, _i, _i, _vp],
    "fz_ccl_areas": [_vp, _vp, _vp, _i, _i, _vp],
    "fz_ccl_table": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_trace_rings": [_vp, _i, _i, _vp, _i, ctypes.c_double, _vp, _vp],
    "fz_trace_rings_fetch": [_vp, _vp, _vp, _vp],
    "fz_onehot_argmax": [_vp, _vp, _i, _i, _i, _i, _vp],
    "fz_ce_workspace_doubles": [_i, _i, _i],
    "fz_ce_loss_forward": [_vp, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "fz_ce_loss_backward": [_vp, _vp, _vp, ctypes.c_float, _vp, _vp, ctypes.c_float, _vp, _i, _i, _i, _i, _vp],
    "fz_confusion_matrix": [_vp, _vp, _i64, _i, _vp, _vp],
    "fz_adamw_step": [_vp, _vp, _vp, _vp, _i64, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                      ctypes.c_double, _i, _vp],
    "fz_adamw_step_dev": [_vp, _vp, _vp, _vp, _i64, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                          ctypes.c_double, _vp, _vp, _vp],
    "fz_transpose_bf16": [_vp, _vp, _i, _i, _vp],
    "fz_colsum_bf16": [_vp, _vp, _vp, _i64, _i, _i, _vp],
    "fz_dwconv7_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_dwconv7_f32_add": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_dwconv7_wgrad": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "fz_layernorm_fwd_stats": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, ctypes.c_float, _i, _vp],
    "fz_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp],
    "fz_gelu_fwd": [_vp, _vp, _i64, _vp],
    "fz_sample_colreduce": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "fz_sample_colreduce2": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "fz_gelu_fwd_sumsq": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "fz_grn_train_forward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, ctypes.c_float, _i, _vp],
    "fz_grn_gelu_backward": [_vp] * 14 + [_i, _i, _i, ctypes.c_float, _vp],
    "fz_grn_gelu_backward_db": [_vp] * 15 + [_i, _i, _i, ctypes.c_float, _vp],
    "fz_grn_gelu_backward_saved": [_vp] * 15 + [_i, _i, _i, ctypes.c_float, _i, _vp],
    "fz_cast_f16_bf16": [_vp, _vp, _i64, _vp],
    "fz_add_f32": [_vp, _vp, _vp, _i64, _vp],
    "fz_reduce_rows_f32": [_vp, _vp, _i, _i, _vp],
    "fz_layernorm_fwd_stats2": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, ctypes.c_float, _i, _vp],
    "fz_s2d_bf16": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "fz_patchify4_nchw": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_im2col3x3_bf16": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_conv3x3_small_supported": [_i, _i, _i, _i],
    "fz_conv3x3_small_forward": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "fz_conv3x3_small_wgrad": [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "fz_col2im3x3": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_bn_relu_train_forward": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, ctypes.c_float, _i, _vp],
    "fz_bn_relu_backward": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i64, _i, _i, _vp],
    "fz_bn_update_running": [_vp, _vp, _vp, _vp, _i, _i64, ctypes.c_float, ctypes.c_float, _vp],
    "fz_upsample2_concat_backward": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fz_head_upsample4": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
}
_RESTYPES = {"fz_last_error": ctypes.c_char_p, "fz_ce_workspace_doubles": ctypes.c_int64}


def exported_symbols():
    """Names the header declares; tests check the library exports every one of them."""
    return list(_SIGNATURES.keys())


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise NativeError(
                f"{_LIB_PATH} is missing: build it with `python -m flair_for_aigle_b200.build` "
                "(there is no CPU fallback for the hot path)")
        _lib = ctypes.CDLL(str(_LIB_PATH))
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        if _lib.fz_abi_version() != ABI_VERSION:
            raise NativeError(f"{_LIB_PATH.name} ABI version mismatch: rebuild with `python -m flair_for_aigle_b200.build`")
    return _lib


def op_dtype() -> torch.dtype:
    """torch dtype of the inference kernels' 16-bit operand format (FZ_OP16 of the header): float16 in the product
    build, bfloat16 in the FZ_OPERANDS=bf16 A/B build.  Training tensors are bfloat16 in both."""
    return torch.float16 if lib().fz_operand_format() == F16 else torch.bfloat16


def _op16(*tensors) -> None:
    want = op_dtype()
    for t in tensors:
        if t is not None and t.dtype != want:
            raise NativeError(f"this kernel takes {want} operands (the library's inference operand format), got {t.dtype}")


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().fz_last_error()
        raise NativeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("native kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise NativeError("native kernels need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise NativeError(f"unsupported dtype {t.dtype}")


# --------------------------------------------------------------------------- feeder
def gather_tiles_f32(raster: torch.Tensor, origins: torch.Tensor, P: int, mean: torch.Tensor, std: torch.Tensor,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    C, H, W = raster.shape
    n = origins.shape[0]
    if raster.dtype not in (torch.uint8, torch.float32) or origins.dtype != torch.int32:
        raise NativeError("gather_tiles_f32: uint8 or float32 raster and int32 origins required")
    if out is None:
        out = torch.empty((n, C, P, P), dtype=torch.float32, device=raster.device)
    fn = lib().fz_gather_tiles_f32 if raster.dtype == torch.uint8 else lib().fz_gather_tiles_f32_from_f32
    _check(fn(_ptr(raster), C, H, W, _ptr(origins), n, P, _ptr(mean), _ptr(std), _ptr(out), _stream()),
           "fz_gather_tiles_f32")
    return out


def gather_tiles_resampled(raster: torch.Tensor, windows: torch.Tensor, ps: int, mean: torch.Tensor, std: torch.Tensor,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rasterio's resampled (bilinear, boundless, fill 0) window read + normalisation: raster uint8 / float32 (C,H,W),
    windows float64 [n,4] = (row_off, col_off, height, width) in the raster's pixels -> float32 [n,C,ps,ps]."""
    C, H, W = raster.shape
    n = windows.shape[0]
    if raster.dtype not in (torch.uint8, torch.float32) or windows.dtype != torch.float64 or windows.shape[1] != 4:
        raise NativeError("gather_tiles_resampled: uint8 / float32 raster and float64 [n,4] windows required")
    if out is None:
        out = torch.empty((n, C, ps, ps), dtype=torch.float32, device=raster.device)
    ratio = float(windows[:, 2:].max().item()) / ps if n else 1.0
    _check(lib().fz_gather_tiles_resampled(_ptr(raster), 1 if raster.dtype == torch.float32 else 0, C, H, W, _ptr(windows), n, ps,
                                           ratio, _ptr(mean), _ptr(std), _ptr(out), _stream()), "fz_gather_tiles_resampled")
    return out


def gather_tiles_u8(raster: torch.Tensor, origins: torch.Tensor, P: int,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    C, H, W = raster.shape
    n = origins.shape[0]
    if out is None:
        out = torch.empty((n, P, P, 4), dtype=torch.uint8, device=raster.device)
    with _Timed('gather_u8', n=n):
        _check(lib().fz_gather_tiles_u8(_ptr(raster), C, H, W, _ptr(origins), n, P, _ptr(out), _stream()),
               "fz_gather_tiles_u8")
    return out


# --------------------------------------------------------------------------- post-processing
def _logits_geom(logits: torch.Tensor, layout: int, n_cls: Optional[int]):
    if layout == NCHW:
        n, c, p, _ = logits.shape
        return n, (n_cls or c), p, 0
    n, p, _, cs = logits.shape
    if layout == NHWC_UP4:
        p *= 4
    return n, (n_cls or cs), p, cs


def crop_argmax_write(logits, layout, margin, plan, own, out_raster, n_cls=None):
    n, c, p, cs = _logits_geom(logits, layout, n_cls)
    H, W = out_raster.shape[-2:]
    _check(lib().fz_crop_argmax_write(_ptr(logits), _dt(logits), layout, cs, n, c, p, margin, _ptr(plan), _ptr(own),
                                      _ptr(out_raster), H, W, _stream()), "fz_crop_argmax_write")


def crop_softmax_write(logits, layout, margin, plan, own, out_raster, n_cls=None):
    n, c, p, cs = _logits_geom(logits, layout, n_cls)
    H, W = out_raster.shape[-2:]
    _check(lib().fz_crop_softmax_write(_ptr(logits), _dt(logits), layout, cs, n, c, p, margin, _ptr(plan), _ptr(own),
                                       _ptr(out_raster), H, W, _stream()), "fz_crop_softmax_write")


def crop_zoom_write(mode: int, logits, layout, margin, plan, own, zmap, out_raster, n_cls=None):
    """argmax (mode 0) / class_prob (mode 1) write with the nearest-neighbour zoom of inference.py:212-226."""
    n, c, p, cs = _logits_geom(logits, layout, n_cls)
    H, W = out_raster.shape[-2:]
    _check(lib().fz_crop_zoom_write(mode, _ptr(logits), _dt(logits), layout, cs, n, c, p, margin, _ptr(plan), _ptr(own),
                                    _ptr(zmap), int(zmap.numel()), _ptr(out_raster), H, W, _stream()),
           "fz_crop_zoom_write")


def _host_plan(plan_host, n):
    """int32 [n,6] numpy copy of the plan for the host-side levelling of overlapping windows (None = one launch per tile)."""
    if plan_host is None:
        return None, None
    import numpy as np
    arr = np.ascontiguousarray(np.asarray(plan_host, dtype=np.int32).reshape(n, 6))
    return arr, ctypes.c_void_p(arr.ctypes.data)


def crop_softmax_accumulate(logits, layout, margin, plan, weight, canvas, n_cls=None, plan_host=None):
    """plan_host: the same plan as a numpy array; with it, tiles whose windows do not overlap share a launch."""
    n, c, p, cs = _logits_geom(logits, layout, n_cls)
    H, W = canvas.shape[-2:]
    keep, php = _host_plan(plan_host, n)
    _check(lib().fz_crop_softmax_accumulate(_ptr(logits), _dt(logits), layout, cs, n, c, p, margin, _ptr(plan), php,
                                            _ptr(weight), _ptr(canvas), H, W, _stream()),
           "fz_crop_softmax_accumulate")


def crop_zoom_accumulate(logits, layout, margin, plan, zmap, canvas, n_cls=None, plan_host=None):
    """canvas += softmax(zoomed cropped logits): the accumulating variant on the rescaled grid (inference.py:515-562)."""
    n, c, p, cs = _logits_geom(logits, layout, n_cls)
    H, W = canvas.shape[-2:]
    keep, php = _host_plan(plan_host, n)
    _check(lib().fz_crop_zoom_accumulate(_ptr(logits), _dt(logits), layout, cs, n, c, p, margin, _ptr(plan), php, _ptr(zmap),
                                         int(zmap.numel()), _ptr(canvas), H, W, _stream()),
           "fz_crop_zoom_accumulate")


def canvas_argmax(canvas: torch.Tensor, want_confidence: bool = False):
    c, H, W = canvas.shape
    labels = torch.empty((H, W), dtype=torch.uint8, device=canvas.device)
    conf = torch.empty((H, W), dtype=torch.float32, device=canvas.device) if want_confidence else None
    _check(lib().fz_canvas_argmax(_ptr(canvas), c, H * W, _ptr(labels), _ptr(conf), _stream()), "fz_canvas_argmax")
    return labels, conf


def convert(img: torch.Tensor, mode: int) -> torch.Tensor:
    C, h, w = img.shape
    out = torch.empty((1, h, w) if mode == 0 else (C, h, w), dtype=torch.uint8, device=img.device)
    _check(lib().fz_convert(_ptr(img), C, h, w, mode, _ptr(out), _stream()), "fz_convert")
    return out


# --------------------------------------------------------------------------- GEMM
def gemm_bf16(A: torch.Tensor, B: torch.Tensor, mode: int, bias=None, resid=None, sumsq=None, out=None,
              rows_per_sample: int = 0, impl: str = "tcgen05") -> torch.Tensor:
    """A: 16-bit [M,K]; B: 16-bit [N,K] or [b,N,K], both float16 (inference) or both bfloat16 (training); a 16-bit
    output has the operands' format.  The format travels to the kernel as FZ_EPI_OPERANDS_F16 in ``mode``."""
    M, K = A.shape
    if B.dim() == 2:
        b_batch, (N, K2) = 1, B.shape
    else:
        b_batch, N, K2 = B.shape
    if K2 != K or A.dtype != B.dtype or A.dtype not in (torch.bfloat16, torch.float16):
        raise NativeError(f"gemm: operands must both be float16 or both bfloat16 with equal K (got {A.dtype} [{M},{K}], "
                          f"{B.dtype} [..,{N},{K2}])")
    f32_out = (mode & 0xff) in (EPI_RESID_F32, EPI_F32)
    if A.dtype == torch.float16:
        mode |= EPI_OPERANDS_F16
    if bias is None:
        bias = torch.zeros(N, dtype=torch.float32, device=A.device)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32 if f32_out else A.dtype, device=A.device)
    elif out.dtype != (torch.float32 if f32_out else A.dtype):
        raise NativeError(f"gemm: output dtype {out.dtype} does not match the epilogue / operand format")
    fn = lib().fz_gemm_bf16 if impl == "tcgen05" else lib().fz_gemm_bf16_simt
    with _Timed("gemm_tcgen05" if impl == "tcgen05" else "gemm_simt", M=M, N=N, K=K, mode=mode):
        _check(fn(_ptr(A), _ptr(B), _ptr(out), _ptr(bias), _ptr(resid), _ptr(sumsq), M, N, K, b_batch,
                  rows_per_sample, mode, _stream()), "fz_gemm_bf16")
    return out


_SPLITK_WS = {}


def gemm_splitk(A: torch.Tensor, B: torch.Tensor, splits: Optional[int] = None, out: Optional[torch.Tensor] = None):
    """fp32 [M,N] = A [M,K] B [N,K]^T for a small output and a long reduction (weight gradients): the K range is split over
    the SMs inside one launch, partial tiles are added in a fixed order.  ``splits`` None = chosen from the shape."""
    M, K = A.shape
    N, K2 = B.shape
    if K2 != K or A.dtype != B.dtype or A.dtype not in (torch.bfloat16, torch.float16):
        raise NativeError("gemm_splitk: operands must both be float16 or both bfloat16 with equal K")
    if splits is None:
        splits = lib().fz_gemm_splitk_max_splits(M, N, K)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    ws = None
    if splits > 1:
        key = (A.device.index, torch.cuda.current_stream().cuda_stream)
        need = splits * M * N
        ws = _SPLITK_WS.get(key)
        if ws is None or ws.numel() < need:
            ws = _SPLITK_WS[key] = torch.empty(need, dtype=torch.float32, device=A.device)
    with _Timed("gemm_splitk", M=M, N=N, K=K, splits=splits):
        _check(lib().fz_gemm_bf16_splitk(_ptr(A), _ptr(B), _ptr(out), _ptr(ws), M, N, K, splits,
                                         EPI_OPERANDS_F16 if A.dtype == torch.float16 else 0, _stream()), "fz_gemm_bf16_splitk")
    return out


def gemm_splitk_tn(At: torch.Tensor, Bt: torch.Tensor, splits: Optional[int] = None, out: Optional[torch.Tensor] = None):
    """fp32 [M,N] = At^T Bt with At [K,M] and Bt [K,N] (both row-major, i.e. transposed operands read in place: MN-major
    tcgen05 descriptors).  dW = dY^T X without transposed copies.  K % 64 == 0, M % 8 == 0, N % 64 == 0."""
    K, M = At.shape
    K2, N = Bt.shape
    if K2 != K or At.dtype != Bt.dtype or At.dtype not in (torch.bfloat16, torch.float16):
        raise NativeError("gemm_splitk_tn: operands must both be float16 or both bfloat16 with equal K")
    if splits is None:
        splits = lib().fz_gemm_splitk_max_splits(M, N, K)
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=At.device)
    ws = None
    if splits > 1:
        key = (At.device.index, torch.cuda.current_stream().cuda_stream)
        need = splits * M * N
        ws = _SPLITK_WS.get(key)
        if ws is None or ws.numel() < need:
            ws = _SPLITK_WS[key] = torch.empty(need, dtype=torch.float32, device=At.device)
    with _Timed("gemm_splitk_tn", M=M, N=N, K=K, splits=splits):
        _check(lib().fz_gemm_bf16_splitk_tn(_ptr(At), _ptr(Bt), _ptr(out), _ptr(ws), M, N, K, splits,
                                            EPI_OPERANDS_F16 if At.dtype == torch.float16 else 0, _stream()),
               "fz_gemm_bf16_splitk_tn")
    return out


# --------------------------------------------------------------------------- ConvNeXt-V2 / U-Net ops
def stem_ln(tiles_u8, w, bias, ln_w, ln_b, out, eps=1e-6):
    B, P = tiles_u8.shape[0], tiles_u8.shape[1]
    C0 = w.shape[1]
    with _Timed('stem_ln', n=B):
        _check(lib().fz_stem_ln(_ptr(tiles_u8), _ptr(w), _ptr(bias), _ptr(ln_w), _ptr(ln_b), _ptr(out), B, P, C0, eps,
                                _stream()), "fz_stem_ln")
    return out


def stem_ln_f32(x_nchw, w, bias, ln_w, ln_b, out, eps=1e-6):
    B, Cin, P, _ = x_nchw.shape
    C0 = w.shape[1]
    _check(lib().fz_stem_ln_f32(_ptr(x_nchw), Cin, _ptr(w), _ptr(bias), _ptr(ln_w), _ptr(ln_b), _ptr(out), B, P, C0,
                                eps, _stream()), "fz_stem_ln_f32")
    return out


def dwconv7_ln(x, wdw, bdw, ln_w, ln_b, out, eps=1e-6):
    B, H, W, C = x.shape
    _op16(out)
    with _Timed('dwconv7_ln', B=B, H=H, C=C):
        _check(lib().fz_dwconv7_ln(_ptr(x), _ptr(wdw), _ptr(bdw), _ptr(ln_w), _ptr(ln_b), _ptr(out), B, H, W, C, eps,
                                   _stream()), "fz_dwconv7_ln")
    return out


def ln2d_s2d(x, ln_w, ln_b, out, eps=1e-6, copy=None):
    """LayerNorm2d + space-to-depth; ``copy`` (bf16 [B,H,W,C]) additionally receives the un-normalised input."""
    B, H, W, C = x.shape
    _op16(out, copy)
    with _Timed('ln2d_s2d', B=B, H=H, C=C):
        _check(lib().fz_ln2d_s2d_copy(_ptr(x), _ptr(ln_w), _ptr(ln_b), _ptr(out), _ptr(copy), B, H, W, C, eps,
                                      _stream()), "fz_ln2d_s2d_copy")
    return out


def grn_scale(partial, tiles_per_sample, gamma, scale, eps=1e-6, scratch=None):
    """partial: f32 [B*tiles_per_sample, K] (fc1 epilogue); scale: f32 [B, K]; scratch: f32 [B*K/64]."""
    B, K = scale.shape
    if scratch is None:
        scratch = torch.empty(B * K // 64, dtype=torch.float32, device=scale.device)
    with _Timed('grn_scale', B=B, K=K):
        _check(lib().fz_grn_scale(_ptr(partial), tiles_per_sample, _ptr(gamma), _ptr(scale), _ptr(scratch), B, K, eps,
                                  _stream()), "fz_grn_scale")
    return scale


def scale_weights(w, scale, out):
    N, K = w.shape
    B = scale.shape[0]
    _op16(w, out)
    with _Timed('scale_weights', B=B, N=N, K=K):
        _check(lib().fz_scale_weights(_ptr(w), _ptr(scale), _ptr(out), B, N, K, _stream()), "fz_scale_weights")
    return out


def scale_rows(h, scale, rows_per_sample):
    M, K = h.shape
    _op16(h)
    with _Timed('scale_rows', M=M, K=K):
        _check(lib().fz_scale_rows(_ptr(h), _ptr(scale), M, K, rows_per_sample, _stream()), "fz_scale_rows")
    return h


def upsample2_concat(a, s, out):
    B, H, W, CT = out.shape
    C1 = a.shape[-1]
    C2 = 0 if s is None else s.shape[-1]
    assert C1 + C2 == CT
    with _Timed('upsample2_concat', B=B, H=H, C=CT):
        _check(lib().fz_upsample2_concat(_ptr(a), _dt(a), _ptr(s), _dt(s) if s is not None else _dt(out), _ptr(out),
                                         _dt(out), B, H, W, C1, C2, _stream()), "fz_upsample2_concat")
    return out


def conv3x3(x, w, scale, bias, mode, out=None, cout=None, cstride=0, plan=None, own=None, raster=None, margin=0,
            stride=1, resid=None):
    """x bf16 [B,Hin,Win,Cin]; w bf16 [rows,3,3,Cin]; scale/bias f32 [rows] (scale may be None); the output is
    [B,Hin/stride,Win/stride,cout]; resid (bf16, output shape) is added before the ReLU in CONV_ADD_RELU_BF16."""
    B, Hin, Win, Cin = x.shape
    H, W = Hin // stride, Win // stride
    rows = w.shape[0]
    cout = rows if cout is None else cout
    RH, RW = (raster.shape[-2], raster.shape[-1]) if raster is not None else (0, 0)
    _op16(x, w, resid, out if mode in (CONV_RELU_BF16, CONV_ADD_RELU_BF16, CONV_BF16) else None)
    with _Timed('conv3x3_tcgen05', B=B, H=H, Cin=Cin, Cout=cout, mode=mode):
        _check(lib().fz_conv3x3_ex(_ptr(x), _ptr(w), _ptr(scale), _ptr(bias), _ptr(out), _ptr(resid), B, H, W, Cin, cout,
                                   rows, stride, mode, cstride, _ptr(plan), _ptr(own), _ptr(raster), RH, RW, margin,
                                   _stream()), "fz_conv3x3_ex")
    return out


def merge_upconv_weights(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3] conv weight -> [Cout, 16, Cin] merged sub-pixel taps of conv3x3(nearest_up2(.)) (fp32 sums),
    tile order ((py*2+px)*2+ra)*2+ca as documented at fz_upconv3x3_bn_relu."""
    groups = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}
    w = w.float()
    tiles = []
    for py in range(2):
        for px in range(2):
            for ra in range(2):
                for ca in range(2):
                    acc = torch.zeros_like(w[:, :, 0, 0])
                    for ky in groups[(py, ra)]:
                        for kx in groups[(px, ca)]:
                            acc = acc + w[:, :, ky, kx]
                    tiles.append(acc)
    return torch.stack(tiles, dim=1)


def upconv3x3_bn_relu(x, w16, scale, bias, out):
    """x bf16 [B,H,W,Cin] -> out bf16 [B,2H,2W,Cout] = relu(bn(conv3x3(nearest_up2(x)))); w16 bf16 [rows,16,Cin]."""
    B, H, W, Cin = x.shape
    cout = out.shape[-1]
    _op16(x, w16, out)
    with _Timed('upconv3x3_tcgen05', B=B, H=H, Cin=Cin, Cout=cout):
        _check(lib().fz_upconv3x3_bn_relu(_ptr(x), _ptr(w16), _ptr(scale), _ptr(bias), _ptr(out), B, H, W, Cin, cout,
                                          w16.shape[0], _stream()), "fz_upconv3x3_bn_relu")
    return out


def catconv3x3_bn_relu(a, skip, w16a, w, scale, bias, out):
    """a bf16 [B,Hs,Ws,C1], skip bf16 [B,2Hs,2Ws,C2] -> out bf16 [B,2Hs,2Ws,Cout] =
    relu(bn(conv3x3(cat(nearest_up2(a), skip)))); w16a = merged taps of the first C1 channels, w = the conv's weights."""
    B, Hs, Ws, C1 = a.shape
    C2 = skip.shape[-1]
    cout = out.shape[-1]
    fl_equiv = 2.0 * B * 4 * Hs * Ws * (4 * C1 + 9 * C2) * cout
    _op16(a, skip, w16a, w, out)
    with _Timed('catconv3x3_tcgen05', B=B, H=2 * Hs, C1=C1, C2=C2, Cout=cout, flop=fl_equiv):
        _check(lib().fz_catconv3x3_bn_relu(_ptr(a), _ptr(skip), _ptr(w16a), _ptr(w), _ptr(scale), _ptr(bias), _ptr(out), B,
                                           Hs, Ws, C1, C2, cout, w.shape[0], _stream()), "fz_catconv3x3_bn_relu")
    return out


def conv7x7s2_bn_relu(x, w, scale, bias, out):
    """x: uint8 [B,P,P,4] or float32 [B,Cin,P,P]; out bf16 [B,P/2,P/2,64]."""
    if x.dtype == torch.uint8:
        B, P, is_f32, cin = x.shape[0], x.shape[1], 0, 4
    else:
        B, cin, P, is_f32 = x.shape[0], x.shape[1], x.shape[2], 1
    _op16(out)
    with _Timed('conv7x7s2', B=B, P=P):
        _check(lib().fz_conv7x7s2_bn_relu(_ptr(x), is_f32, cin, _ptr(w), _ptr(scale), _ptr(bias), _ptr(out), B, P,
                                          _stream()), "fz_conv7x7s2_bn_relu")
    return out


def maxpool3x3s2(x, out):
    B, H, W, C = x.shape
    _op16(x, out)
    with _Timed('maxpool3x3s2', B=B, H=H, C=C):
        _check(lib().fz_maxpool3x3s2(_ptr(x), _ptr(out), B, H, W, C, _stream()), "fz_maxpool3x3s2")
    return out


# --------------------------------------------------------------------------- Swin / UPerNet ops
def layernorm_rows(x: torch.Tensor, w, b, out: torch.Tensor, eps: float = 1e-5):
    """x float [..., C] (contiguous) -> out bf16, nn.LayerNorm over C."""
    C = x.shape[-1]
    rows = x.numel() // C
    _op16(out)
    with _Timed("layernorm_rows", rows=rows, C=C):
        _check(lib().fz_layernorm_rows(_ptr(x), _ptr(w), _ptr(b), _ptr(out), rows, C, eps, _stream()), "fz_layernorm_rows")
    return out


def merge_ln(x: torch.Tensor, w, b, out: torch.Tensor, eps: float = 1e-5):
    """timm PatchMerging gather + LayerNorm(4C): x float [B,H,W,C] -> out bf16 [B,H/2,W/2,4C]."""
    B, H, W, C = x.shape
    _op16(out)
    with _Timed("merge_ln", n=B, H=H, C=C):
        _check(lib().fz_merge_ln(_ptr(x), _ptr(w), _ptr(b), _ptr(out), B, H, W, C, eps, _stream()), "fz_merge_ln")
    return out


def swin_window_attn(qkv: torch.Tensor, qkv_bias_bf16, table, out: torch.Tensor, heads: int, window: int, shift: int,
                     scale: float):
    """qkv bf16 [B,H,W,3C] -> out bf16 [B,H,W,C] (see include/flair_zonal_b200.h)."""
    B, H, W, C3 = qkv.shape
    _op16(qkv, qkv_bias_bf16, out)
    with _Timed("swin_window_attn", n=B, H=H, C=C3 // 3):
        _check(lib().fz_swin_window_attn(_ptr(qkv), _ptr(qkv_bias_bf16), _ptr(table), _ptr(out), B, H, W, C3 // 3, heads,
                                         window, shift, scale, _stream()), "fz_swin_window_attn")
    return out


def cast_f32_bf16(x: torch.Tensor, out: torch.Tensor):
    """fp32 -> ``out``'s 16-bit format (float16: saturating; bfloat16), round to nearest even."""
    with _Timed("cast_f32_bf16", n=x.numel()):
        _check(lib().fz_cast_f32_16(_ptr(x), _ptr(out), _dt(out), x.numel(), _stream()), "fz_cast_f32_16")
    return out


def adaptive_avgpool(x: torch.Tensor, S: int, out: torch.Tensor):
    B, H, W, C = x.shape
    _op16(x, out)
    with _Timed("adaptive_avgpool", n=B, S=S):
        _check(lib().fz_adaptive_avgpool(_ptr(x), _ptr(out), B, H, W, C, S, _stream()), "fz_adaptive_avgpool")
    return out


def bilinear_slice(x: torch.Tensor, out: torch.Tensor, c0: int = 0, add=None):
    """out[..., c0:c0+C] = bilinear(x -> out's H,W; align_corners=False) (+ add)."""
    B, h, w, C = x.shape
    _, H, W, Ctot = out.shape
    _op16(x, add, out)
    with _Timed("bilinear_slice", n=B, H=H, C=C):
        _check(lib().fz_bilinear_slice(_ptr(x), _ptr(add), _ptr(out), B, h, w, H, W, C, Ctot, c0, _stream()),
               "fz_bilinear_slice")
    return out


def updown_slice(x: torch.Tensor, out: torch.Tensor, c0: int = 0):
    B, H, W, C = x.shape
    Ctot = out.shape[-1]
    _op16(x, out)
    with _Timed("updown_slice", n=B, H=H, C=C):
        _check(lib().fz_updown_slice(_ptr(x), _ptr(out), B, H, W, C, Ctot, c0, _stream()), "fz_updown_slice")
    return out


def pyramid_concat(p0, p1, p2, p3, out: torch.Tensor):
    """out [B,H,H,5C] = [bilinear(p0) | bilinear(p1) | bilinear(p2) | p3 | down2(up2(p3))]: the UPerNet fuse input."""
    B, H, _, C = p3.shape
    for t in (p0, p1, p2, p3, out):
        if t.dtype != op_dtype() or not t.is_contiguous():
            raise NativeError("pyramid_concat: contiguous NHWC maps in the inference operand format required")
    if tuple(out.shape) != (B, H, H, 5 * C):
        raise NativeError(f"pyramid_concat: out shape {tuple(out.shape)} != {(B, H, H, 5 * C)}")
    with _Timed("pyramid_concat", n=B, H=H, C=C):
        _check(lib().fz_pyramid_concat(_ptr(p0), p0.shape[1], _ptr(p1), p1.shape[1], _ptr(p2), p2.shape[1], _ptr(p3),
                                       _ptr(out), B, H, C, _stream()), "fz_pyramid_concat")
    return out


def head_upsample4(logits: torch.Tensor, n_cls: int, out: torch.Tensor):
    """logits float [B,h,w,cstride] -> out float [B,n_cls,4h,4w] (UpsamplingBilinear2d, align_corners=True)."""
    B, h, w, cs = logits.shape
    with _Timed("head_upsample4", n=B):
        _check(lib().fz_head_upsample4(_ptr(logits), _ptr(out), B, h, w, cs, n_cls, _stream()), "fz_head_upsample4")
    return out


# ------------------------------------------------------------------------------------------------ polygonisation
def ccl_label(raster: torch.Tensor) -> torch.Tensor:
    """uint8 class raster [H,W] on the device -> int32 labels [H,W]: 4-connected components of equal class, label =
    smallest linear pixel index of the component (rasterio.features.shapes' connectivity, inference.py:366)."""
    if raster.dtype != torch.uint8 or raster.dim() != 2 or not raster.is_contiguous():
        raise NativeError("ccl_label: contiguous uint8 [H,W] raster required")
    H, W = raster.shape
    labels = torch.empty((H, W), dtype=torch.int32, device=raster.device)
    with _Timed("ccl_label", H=H, W=W):
        _check(lib().fz_ccl_label(_ptr(raster), _ptr(labels), H, W, _stream()), "fz_ccl_label")
    return labels


def ccl_components(raster: torch.Tensor, labels: torch.Tensor, min_area_px: int = 0, ignore_class: int = -1):
    """-> (roots int32[n], areas int32[n], classes uint8[n]) as numpy arrays sorted by root (= by first pixel), for the
    components with at least ``min_area_px`` pixels whose class is not ``ignore_class``; also returns nothing else: the
    total component count is ``ccl_components.last_total``."""
    import numpy as np
    H, W = raster.shape
    area = torch.zeros(H * W, dtype=torch.int32, device=raster.device)
    counter = torch.zeros(3, dtype=torch.int32, device=raster.device)
    _check(lib().fz_ccl_areas(_ptr(labels), _ptr(area), _ptr(counter), H, W, _stream()), "fz_ccl_areas")
    cap = 1 << 16
    while True:
        rec = torch.empty((cap, 3), dtype=torch.int32, device=raster.device)
        counter[1:].zero_()
        _check(lib().fz_ccl_table(_ptr(raster), _ptr(labels), _ptr(area), _ptr(counter[1:]), _ptr(rec), cap,
                                  int(min_area_px), int(ignore_class), H, W, _stream()), "fz_ccl_table")
        total, n = (int(v) for v in counter[:2].tolist())
        if n <= cap:
            break
        cap = n
    ccl_components.last_total = total
    rec = rec[:n].cpu().numpy()
    rec = rec[np.argsort(rec[:, 0], kind="stable")]
    return rec[:, 0].copy(), rec[:, 1].copy(), rec[:, 2].astype(np.uint8)


def trace_rings(labels_host, keep_roots, simplify_px: float = 0.0):
    """Host ring tracer on an int32 label image (numpy, C-contiguous).  -> (ring_root int32[r], ring_is_hole bool[r],
    ring_offset int64[r+1], xy float64[p,2]) in pixel-corner coordinates, rings closed."""
    import numpy as np
    labels_host = np.ascontiguousarray(labels_host, dtype=np.int32)
    keep = np.ascontiguousarray(np.sort(np.asarray(keep_roots, dtype=np.int32)))
    H, W = labels_host.shape
    nr, npnt = ctypes.c_int64(0), ctypes.c_int64(0)
    _check(lib().fz_trace_rings(labels_host.ctypes.data, H, W, keep.ctypes.data if keep.size else None, int(keep.size),
                                float(simplify_px), ctypes.addressof(nr), ctypes.addressof(npnt)), "fz_trace_rings")
    ring_root = np.empty(nr.value, np.int32)
    ring_hole = np.empty(nr.value, np.uint8)
    ring_off = np.empty(nr.value + 1, np.int64)
    xy = np.empty((npnt.value, 2), np.float64)
    _check(lib().fz_trace_rings_fetch(ring_root.ctypes.data, ring_hole.ctypes.data, ring_off.ctypes.data, xy.ctypes.data),
           "fz_trace_rings_fetch")
    return ring_root, ring_hole.astype(bool), ring_off, xy


# ------------------------------------------------------------------------------------------------ training: loss, optimizer
def onehot_argmax(onehot: torch.Tensor) -> torch.Tensor:
    """tasks_module.py:154: (B,C,H,W) float one-hot labels -> int32 (B,H,W) class indices."""
    B, C, H, W = onehot.shape
    out = torch.empty((B, H, W), dtype=torch.int32, device=onehot.device)
    _check(lib().fz_onehot_argmax(_ptr(onehot.float().contiguous()), _ptr(out), B, C, H, W, _stream()), "fz_onehot_argmax")
    return out


def confusion_matrix(targets: torch.Tensor, preds: torch.Tensor, num_classes: int, out: Optional[torch.Tensor] = None):
    """int64 [C][C] counts, rows = label, columns = prediction (sklearn / torchmetrics convention); labels or predictions
    outside 0..C-1 are skipped.  ``out``: an existing matrix to accumulate into."""
    if targets.shape != preds.shape:
        raise NativeError(f"confusion_matrix: shapes {tuple(targets.shape)} vs {tuple(preds.shape)}")
    t = targets.to(torch.int32).contiguous()
    p = preds.to(torch.int32).contiguous()
    if out is None:
        out = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=t.device)
    elif out.dtype != torch.int64 or tuple(out.shape) != (num_classes, num_classes):
        raise NativeError("confusion_matrix: out must be int64 [C][C]")
    _check(lib().fz_confusion_matrix(_ptr(t), _ptr(p), t.numel(), int(num_classes), _ptr(out), _stream()), "fz_confusion_matrix")
    return out


def ce_loss_forward(logits: torch.Tensor, targets: torch.Tensor, class_weight: Optional[torch.Tensor],
                    task_weight: float = 1.0, want_preds: bool = True):
    """-> (loss_out float[2] = {task_weight * CE, sum of w[t]}, lse (B,H,W), preds int32 (B,H,W) or None)."""
    if logits.dtype != torch.float32 or logits.dim() != 4:
        raise NativeError("ce_loss_forward: fp32 (B,C,H,W) logits required")
    B, C, H, W = logits.shape
    dev = logits.device
    lse = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    preds = torch.empty((B, H, W), dtype=torch.int32, device=dev) if want_preds else None
    ws = torch.empty(int(lib().fz_ce_workspace_doubles(B, H, W)), dtype=torch.float64, device=dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    with _Timed("ce_loss_forward", n=B, C=C, H=H):
        _check(lib().fz_ce_loss_forward(_ptr(logits), _ptr(targets), _ptr(class_weight), float(task_weight), _ptr(lse),
                                        _ptr(preds), _ptr(ws), _ptr(out), B, C, H, W, _stream()), "fz_ce_loss_forward")
    return out, lse, preds


def ce_loss_backward(logits, targets, class_weight, task_weight, lse, loss_out, grad_scale: float = 1.0, out=None):
    B, C, H, W = logits.shape
    dlogits = out if out is not None else torch.empty_like(logits)
    with _Timed("ce_loss_backward", n=B, C=C, H=H):
        _check(lib().fz_ce_loss_backward(_ptr(logits), _ptr(targets), _ptr(class_weight), float(task_weight), _ptr(lse),
                                         _ptr(loss_out), float(grad_scale), _ptr(dlogits), B, C, H, W, _stream()),
               "fz_ce_loss_backward")
    return dlogits


def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step: int):
    """torch.optim.AdamW's update on flat contiguous fp32 buffers, in place."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or t.numel() != param.numel():
            raise NativeError("adamw_step: fp32 buffers of one size required")
    with _Timed("adamw_step", n=param.numel()):
        _check(lib().fz_adamw_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr),
                                   float(beta1), float(beta2), float(eps), float(weight_decay), int(step), _stream()),
               "fz_adamw_step")


def adamw_step_dev(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step_dev, hyper_dev):
    """The same update with the step counter on the device (``step_dev`` int64[1], incremented by the call; ``hyper_dev``
    float32[2] scratch): no step-dependent host argument, so a captured step replays correctly as a CUDA graph."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or t.numel() != param.numel():
            raise NativeError("adamw_step_dev: fp32 buffers of one size required")
    if step_dev.dtype != torch.int64 or hyper_dev.dtype != torch.float32 or hyper_dev.numel() < 2:
        raise NativeError("adamw_step_dev: step_dev int64[1] and hyper_dev float32[2] required")
    _check(lib().fz_adamw_step_dev(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr),
                                   float(beta1), float(beta2), float(eps), float(weight_decay), _ptr(step_dev),
                                   _ptr(hyper_dev), _stream()), "fz_adamw_step_dev")


def transpose_bf16(x: torch.Tensor) -> torch.Tensor:
    """bf16 [R,C] -> contiguous bf16 [C,R]."""
    if x.dtype != torch.bfloat16 or x.dim() != 2:
        raise NativeError("transpose_bf16: bf16 [R,C] required")
    R, C = x.shape
    out = torch.empty((C, R), dtype=torch.bfloat16, device=x.device)
    with _Timed("transpose_bf16", R=R, C=C):
        _check(lib().fz_transpose_bf16(_ptr(x), _ptr(out), R, C, _stream()), "fz_transpose_bf16")
    return out


def colsum_bf16(x: torch.Tensor, chunks: int = 64) -> torch.Tensor:
    """bf16 [M,N] -> float32 [N] column sums (deterministic two-stage reduction)."""
    M, N = x.shape
    chunks = max(1, min(chunks, M))
    partial = torch.empty((chunks, N), dtype=torch.float32, device=x.device)
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    _check(lib().fz_colsum_bf16(_ptr(x), _ptr(partial), _ptr(out), M, N, chunks, _stream()), "fz_colsum_bf16")
    return out


_STREAMS = {}


def _stream_pool(device, n: int):
    key = (device.index, n)
    if key not in _STREAMS:
        _STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
    return _STREAMS[key]


def weight_gradient(dY: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """fp32 [N,K] = dY^T X for dY [M,N] and X [M,K] (16-bit; same format, or bf16 dY with fp16 X): split-K, operands read in place.  Rows are padded
    to a multiple of 64 (the reduction's k-block) only when needed; shapes the MN-major kernel does not take (N % 8 != 0 or
    K % 64 != 0) go through transposed copies."""
    M, N = dY.shape
    K = X.shape[1]
    if X.dtype != dY.dtype:
        # bf16 gradients x fp16 forward activations: one MMA takes one operand format (a mixed descriptor is an illegal
        # instruction), so the activations are re-rounded to bf16 here -- a benign rounding, the gradient noise that matters
        # comes from the FORWARD values (tests/diag/grad_precision_budget.py)
        if dY.dtype != torch.bfloat16 or X.dtype != torch.float16:
            raise NativeError(f"weight_gradient: formats {dY.dtype} / {X.dtype} (bfloat16 gradients with float16 or bfloat16 activations)")
        Xb = torch.empty(X.shape, dtype=torch.bfloat16, device=X.device)
        _check(lib().fz_cast_f16_bf16(_ptr(X.contiguous()), _ptr(Xb), X.numel(), _stream()), "fz_cast_f16_bf16")
        X = Xb
    if M % 64:
        Mp = (M + 63) // 64 * 64
        dYp = torch.zeros((Mp, N), dtype=dY.dtype, device=dY.device)
        Xp = torch.zeros((Mp, K), dtype=X.dtype, device=X.device)
        dYp[:M].copy_(dY)
        Xp[:M].copy_(X)
        dY, X = dYp, Xp
    if N % 8 == 0 and K % 64 == 0:
        return gemm_splitk_tn(dY, X)
    return gemm_splitk(transpose_bf16(dY), transpose_bf16(X))


def linear_backward(dY: torch.Tensor, X: torch.Tensor, W: torch.Tensor, db: torch.Tensor = None):
    """Gradients of Y = X W^T + b (X bf16 [M,K], W bf16 [N,K], dY bf16 [M,N]) on the tcgen05 GEMM:
    dX = dY W (bf16 [M,K]), dW = dY^T X (fp32 [N,K], fp32 accumulation over all M rows), db = column sums of dY
    (``db``: already computed by the kernel that produced dY -- returned as is)."""
    M, N = dY.shape
    K = X.shape[1]
    if X.shape[0] != M or tuple(W.shape) != (N, K):
        raise NativeError("linear_backward: shape mismatch")
    dX = gemm_bf16(dY, transpose_bf16(W), EPI_BF16)                               # [M,N] x [K,N]^T
    # dW = dY^T X: a small [N,K] output reduced over all M rows.  Split-K inside the GEMM kernel (round 1 ran 2..32 CTAs per
    # launch, or row chunks on side streams), and both operands are read IN PLACE through MN-major tcgen05 descriptors
    # (round 1 made transposed copies of dY and X for every layer: 499 launches, 19 ms per step)
    dW = weight_gradient(dY, X)
    return dX, dW, (colsum_bf16(dY) if db is None else db)

"""ConvNeXt-V2 encoder + U-Net decoder forward, driven kernel by kernel through the C ABI.

This is the B200 execution plan for the model the reference builds with
``smp.create_model(arch='unet', encoder_name='tu-convnextv2_*')`` and runs at
``flair_hub/models/flair_model.py:376`` (encoder) and ``:417-419`` (decoder + head).  It owns
the packed weights (bf16 GEMM operands in the K-major layouts the tcgen05 kernels read, fp32
vectors for the CUDA-core kernels) and a workspace sized for ``max_batch`` tiles; PyTorch only
provides the device buffers and the stream.

Numerics: GEMM / conv operands bf16, fp32 accumulation in TMEM, fp32 residual stream, fp32
LayerNorm / GRN statistics, eval-mode BatchNorm applied as an fp32 per-channel scale + bias in
the conv epilogue.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from .. import native as nv
from .unet_decoder import UNetDecoderPlan


@dataclass
class ConvNeXtCfg:
    depths: Sequence[int] = (3, 3, 27, 3)
    dims: Sequence[int] = (128, 256, 512, 1024)
    in_chans: int = 4
    decoder_channels: Sequence[int] = (256, 128, 64, 32, 16)
    n_classes: int = 19
    patch: int = 512


CONVNEXTV2_CFGS = {
    "convnextv2_atto": ((2, 2, 6, 2), (40, 80, 160, 320)),
    "convnextv2_femto": ((2, 2, 6, 2), (48, 96, 192, 384)),
    "convnextv2_pico": ((2, 2, 6, 2), (64, 128, 256, 512)),
    "convnextv2_nano": ((2, 2, 8, 2), (80, 160, 320, 640)),
    "convnextv2_tiny": ((3, 3, 9, 3), (96, 192, 384, 768)),
    "convnextv2_base": ((3, 3, 27, 3), (128, 256, 512, 1024)),
    "convnextv2_large": ((3, 3, 27, 3), (192, 384, 768, 1536)),
    "convnextv2_huge": ((3, 3, 27, 3), (352, 704, 1408, 2816)),
}


def _f32(t, dev):
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _bf16(t, dev):
    """fp32 -> the inference kernels' 16-bit operand format (nv.op_dtype(): float16, saturating; bfloat16 in the A/B
    build), one rounding."""
    t = t.detach().to(device=dev, dtype=torch.float32)
    dt = nv.op_dtype()
    if dt == torch.float16:
        t = t.clamp(-65504.0, 65504.0)
    return t.to(dt).contiguous()


class ConvNeXtV2UNetEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], enc_prefix: str, dec_prefix: str, cfg: ConvNeXtCfg,
                 device: torch.device, max_batch: int = 16, norm_mean: Optional[Sequence[float]] = None,
                 norm_std: Optional[Sequence[float]] = None, bn_eps: float = 1e-5):
        if device.type != "cuda":
            raise nv.NativeError("ConvNeXtV2UNetEngine needs a CUDA device (no CPU fallback)")
        for d in cfg.dims:
            if d % 128 != 0:
                raise NotImplementedError(f"encoder width {d}: the sm_100a kernels are tiled for multiples of 128")
        if cfg.patch % 512 != 0:
            # the fc1 epilogue emits GRN partial sums per 128-row tile and tiles must not straddle samples:
            # (patch/32)^2 rows per sample in the last stage have to be a multiple of 128
            raise NotImplementedError(f"patch size {cfg.patch}: the ConvNeXt-V2 plan needs a multiple of 512")
        nv.lib()
        self.cfg, self.dev, self.B = cfg, device, max_batch
        self.gemm_impl = "tcgen05"
        sd = {k: v.detach().to('cpu') for k, v in state_dict.items()}   # pack on the host, upload once
        E, D, dev = enc_prefix, dec_prefix, device
        C0 = cfg.dims[0]

        # ---- stem: [C0, Cin, 4, 4] -> [64][C0], k = ky*16 + kx*4 + c (c padded to 4)
        w = sd[E + "stem_0.weight"].float()
        assert w.shape == (C0, cfg.in_chans, 4, 4), w.shape
        wk = torch.zeros(4, 4, 4, C0, dtype=torch.float64)
        wk[:, :, :cfg.in_chans, :] = w.double().permute(2, 3, 1, 0)
        self.stem_w_f32 = _f32(wk.reshape(64, C0), dev)           # for already-normalised float input
        self.stem_b_f32 = _f32(sd[E + "stem_0.bias"], dev)
        if norm_mean is not None:
            mean = torch.zeros(4, dtype=torch.float64)
            std = torch.ones(4, dtype=torch.float64)
            mean[:cfg.in_chans] = torch.tensor(list(norm_mean), dtype=torch.float64)
            std[:cfg.in_chans] = torch.tensor(list(norm_std), dtype=torch.float64)
            wf = wk / std.view(1, 1, 4, 1)
            bf = sd[E + "stem_0.bias"].double() - (wf * mean.view(1, 1, 4, 1)).sum(dim=(0, 1, 2))
            self.stem_w_u8 = _f32(wf.reshape(64, C0), dev)        # raw uint8 input, normalisation folded
            self.stem_b_u8 = _f32(bf, dev)
        else:
            self.stem_w_u8 = self.stem_b_u8 = None
        self.stem_ln_w = _f32(sd[E + "stem_1.weight"], dev)
        self.stem_ln_b = _f32(sd[E + "stem_1.bias"], dev)

        # ---- stages
        self.stages: List[dict] = []
        for i, (depth, C) in enumerate(zip(cfg.depths, cfg.dims)):
            S = E + f"stages_{i}."
            st = {"C": C, "blocks": []}
            if i > 0:
                Ci = cfg.dims[i - 1]
                st["ds_ln_w"] = _f32(sd[S + "downsample.0.weight"], dev)
                st["ds_ln_b"] = _f32(sd[S + "downsample.0.bias"], dev)
                wd = sd[S + "downsample.1.weight"].float()          # [C, Ci, 2, 2] -> [C][ky][kx][Ci]
                st["ds_w"] = _bf16(wd.permute(0, 2, 3, 1).reshape(C, 4 * Ci), dev)
                st["ds_b"] = _f32(sd[S + "downsample.1.bias"], dev)
            for j in range(depth):
                Bk = S + f"blocks.{j}."
                w2 = sd[Bk + "mlp.fc2.weight"].float()              # [C, 4C]
                beta = sd[Bk + "mlp.grn.bias"].float()
                blk = {
                    "dw_w": _f32(sd[Bk + "conv_dw.weight"].float().reshape(C, 49).t(), dev),   # [49][C]
                    "dw_b": _f32(sd[Bk + "conv_dw.bias"], dev),
                    "ln_w": _f32(sd[Bk + "norm.weight"], dev),
                    "ln_b": _f32(sd[Bk + "norm.bias"], dev),
                    "fc1_w": _bf16(sd[Bk + "mlp.fc1.weight"], dev),                            # [4C, C]
                    "fc1_b": _f32(sd[Bk + "mlp.fc1.bias"], dev),
                    "grn_g": _f32(sd[Bk + "mlp.grn.weight"], dev),
                    "fc2_w": _bf16(w2, dev),                                                   # [C, 4C]
                    # GRN: fc2(x*s + beta) = (W2 diag(s)) x + (W2 beta + b2)
                    "fc2_b": _f32(sd[Bk + "mlp.fc2.bias"].double() + w2.double() @ beta.double(), dev),
                }
                st["blocks"].append(blk)
            self.stages.append(st)

        # ---- decoder + head (smp UnetDecoder / SegmentationHead), shared with the other encoders
        self.decoder = UNetDecoderPlan(sd, D, [cfg.in_chans, 0] + list(cfg.dims), cfg.n_classes, cfg.patch, max_batch,
                                       dev, decoder_channels=cfg.decoder_channels, bn_eps=bn_eps)
        self.dec = self.decoder.blocks
        self._alloc_workspace()

    # ------------------------------------------------------------------------------ workspace
    def _alloc_workspace(self):
        cfg, B, dev = self.cfg, self.B, self.dev
        P = cfg.patch
        bf, f32 = nv.op_dtype(), torch.float32
        hw = [(P // 4) >> i for i in range(4)]
        self.hw = hw
        self.x = [torch.empty((B, h, h, c), dtype=f32, device=dev) for h, c in zip(hw, cfg.dims)]
        # bf16 copies of the stage outputs = the decoder's operands (written by the downsample LayerNorm kernel while
        # the rows are in registers; the last stage by a cast)
        self.xb = [torch.empty((B, h, h, c), dtype=bf, device=dev) for h, c in zip(hw, cfg.dims)]
        self.copy_ok = [c in (128, 256, 512) for c in cfg.dims]
        max_y = max(h * h * c for h, c in zip(hw, cfg.dims))
        self.y = torch.empty(B * max_y, dtype=bf, device=dev)            # dwconv+LN out / s2d operand
        self.h = torch.empty(B * max_y * 4, dtype=bf, device=dev)        # MLP hidden
        kmax = 4 * max(cfg.dims)
        # fc1 epilogue partials: one row of 4C sums per 128-row tile
        self.sumsq = torch.zeros(B * max((h * h // 128) * 4 * c for h, c in zip(hw, cfg.dims)), dtype=f32, device=dev)
        self.scale = torch.empty(B * kmax, dtype=f32, device=dev)
        self.grn_scratch = torch.empty(B * kmax // 64, dtype=f32, device=dev)
        # per-sample GRN-scaled fc2 weights where that is cheaper than scaling the hidden rows
        self.use_wscale = [h * h > c for h, c in zip(hw, cfg.dims)]
        wmax = max([4 * c * c for c, u in zip(cfg.dims, self.use_wscale) if u] + [0])
        self.w2s = torch.empty(B * wmax, dtype=bf, device=dev) if wmax else None

    # ------------------------------------------------------------------------------ encoder
    def _gemm(self, A, Bw, mode, **kw):
        return nv.gemm_bf16(A, Bw, mode, impl=self.gemm_impl, **kw)

    def _encode(self, n: int) -> None:
        """Stages 0..3 on the whole batch.  (Running the blocks over L2-sized sub-batches of tiles, so that the dwconv
        output / hidden tensors never leave the 126 MB L2, was measured slower in every setting -- DESIGN.md section 4 --
        and is gone.)"""
        cfg = self.cfg
        for i, st in enumerate(self.stages):
            C, hwi = st["C"], self.hw[i]
            rps = hwi * hwi
            M = n * rps
            if i > 0:
                Ci = cfg.dims[i - 1]
                s2d = self.y[:M * 4 * Ci].view(M, 4 * Ci)
                nv.ln2d_s2d(self.x[i - 1][:n], st["ds_ln_w"], st["ds_ln_b"], s2d,
                            copy=self.xb[i - 1][:n] if self.copy_ok[i - 1] else None)
                if not self.copy_ok[i - 1]:
                    nv.cast_f32_bf16(self.x[i - 1][:n], self.xb[i - 1][:n])
                self._gemm(s2d, st["ds_w"], nv.EPI_F32, bias=st["ds_b"], out=self.x[i][:n].view(M, C))
            tps = rps // 128
            x = self.x[i][:n]
            xm = x.view(M, C)
            y = self.y[:M * C].view(n, hwi, hwi, C)
            hbuf = self.h[:M * 4 * C].view(M, 4 * C)
            sumsq = self.sumsq[:n * tps * 4 * C].view(n * tps, 4 * C)
            scale = self.scale[:n * 4 * C].view(n, 4 * C)
            for blk in st["blocks"]:
                nv.dwconv7_ln(x, blk["dw_w"], blk["dw_b"], blk["ln_w"], blk["ln_b"], y)
                self._gemm(y.view(M, C), blk["fc1_w"], nv.EPI_GELU_SUMSQ, bias=blk["fc1_b"], sumsq=sumsq, out=hbuf,
                           rows_per_sample=rps)
                nv.grn_scale(sumsq, tps, blk["grn_g"], scale, scratch=self.grn_scratch)
                # fc2 walks its tiles backwards: fc1 has just streamed the hidden tensor out (larger than L2 in
                # stages 0-2), so its newest rows are still cached; it then finishes on the rows the next
                # block's dwconv starts with.  Measured -11 % on the stage-2 fc1+fc2 pair.
                fc2_mode = nv.EPI_RESID_F32 | nv.EPI_REVERSE_TILES
                if self.use_wscale[i]:
                    w2s = self.w2s[:n * 4 * C * C].view(n, C, 4 * C)
                    nv.scale_weights(blk["fc2_w"], scale, w2s)
                    self._gemm(hbuf, w2s, fc2_mode, bias=blk["fc2_b"], resid=xm, out=xm, rows_per_sample=rps)
                else:
                    nv.scale_rows(hbuf, scale, rps)
                    self._gemm(hbuf, blk["fc2_w"], fc2_mode, bias=blk["fc2_b"], resid=xm, out=xm, rows_per_sample=rps)

    def encode_u8(self, tiles_u8: torch.Tensor) -> None:
        if self.stem_w_u8 is None:
            raise nv.NativeError("engine was built without normalisation constants: uint8 input unavailable")
        n = tiles_u8.shape[0]
        assert n <= self.B
        nv.stem_ln(tiles_u8, self.stem_w_u8, self.stem_b_u8, self.stem_ln_w, self.stem_ln_b, self.x[0][:n])
        self._encode(n)

    def encode_f32(self, x_nchw: torch.Tensor) -> None:
        n = x_nchw.shape[0]
        assert n <= self.B and x_nchw.shape[1] == self.cfg.in_chans
        nv.stem_ln_f32(x_nchw, self.stem_w_f32, self.stem_b_f32, self.stem_ln_w, self.stem_ln_b, self.x[0][:n])
        self._encode(n)

    def features(self, n: int) -> List[torch.Tensor]:
        """Stage outputs (strides 4/8/16/32) as fp32 NHWC views."""
        return [x[:n] for x in self.x]

    # ------------------------------------------------------------------------------ decoder
    def _feats_deep_first(self, n: int):
        nv.cast_f32_bf16(self.x[3][:n], self.xb[3][:n])
        return [self.xb[3][:n], self.xb[2][:n], self.xb[1][:n], self.xb[0][:n]]

    def decode_logits_nchw(self, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 logits [n, n_classes, P, P] -- the layout FLAIR_HUB_Model.forward returns."""
        return self.decoder.logits_nchw(self._feats_deep_first(n), n, out)

    def decode_logits_nhwc(self, n: int, out: torch.Tensor) -> torch.Tensor:
        return self.decoder.logits_nhwc(self._feats_deep_first(n), n, out)

    def decode_argmax_to_raster(self, n: int, plan: torch.Tensor, own: Optional[torch.Tensor], raster: torch.Tensor,
                                margin: int) -> None:
        """Head conv with the crop + argmax + last-writer-wins write fused into its epilogue
        (inference.py:295-352): no logits leave the SM."""
        self.decoder.argmax_to_raster(self._feats_deep_first(n), n, plan, own, raster, margin)

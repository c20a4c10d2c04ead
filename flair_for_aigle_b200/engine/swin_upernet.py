"""Swin encoder + UPerNet decoder execution plan -- `swin_base_patch4_window12_384-upernet`, BASELINE.json configs[2].

What the reference builds with ``smp.create_model('upernet', encoder_name='tu-swin_base_patch4_window12_384',
img_size=512)`` (monotemp_model.py:67-92) and runs at flair_model.py:376 (encoder) / :417-419 (decoder + head).
Semantics follow recent timm (restated and pinned in oracle/swin_upernet.py):

  patch embed  conv4x4/s4 + LayerNorm(1e-5)                       csrc/convnext_ops.cu stem kernel -> fp32 NHWC stream
  block        LN -> qkv GEMM -> window attention (roll, pad, rel-pos bias, shift mask) -> proj GEMM (+residual)
               LN -> fc1 GEMM (+GELU) -> fc2 GEMM (+residual)    csrc/swin_ops.cu + tcgen05 GEMMs, fp32 residual stream
  merging      2x2 gather + LN(4C) -> bias-free GEMM              at the START of stages 1..3
  UPerNet      PSP (avg-pool 1/2/3/6 -> 1x1 -> bilinear) -> 1x1; FPN laterals (1x1 GEMMs) + bilinear top-down adds;
               five maps resized to H/4, concat 1280 -> conv3x3 (tcgen05 implicit GEMM) -> 1x1 head GEMM ->
               bilinear x4 (align_corners=True) -> fp32 logits NCHW -> crop/argmax kernels (csrc/postprocess.cu)

Eval BatchNorm of the 1x1 convolutions is folded into their bf16 weights (fp64 product, one rounding) and fp32 bias;
the 3x3 fuse convolution keeps it as an fp32 scale/bias in the epilogue.  Activations bf16 NHWC.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from .. import native as nv
from .unet_decoder import _bf16, _f32

SWIN_CFGS = {
    # name: (embed_dim, depths, heads, window)
    "swin_base_patch4_window7_224": (128, (2, 2, 18, 2), (4, 8, 16, 32), 7),
    "swin_base_patch4_window12_384": (128, (2, 2, 18, 2), (4, 8, 16, 32), 12),
}

PSP_SIZES = (1, 2, 3, 6)


@dataclass
class SwinCfg:
    embed_dim: int = 128
    depths: Sequence[int] = (2, 2, 18, 2)
    heads: Sequence[int] = (4, 8, 16, 32)
    window: int = 12
    in_chans: int = 4
    n_classes: int = 19
    patch: int = 512
    pyramid_channels: int = 256
    segmentation_channels: int = 64


def _bn_fold(sd, key, eps=1e-5):
    g, b = sd[key + ".weight"].double(), sd[key + ".bias"].double()
    mu, var = sd[key + ".running_mean"].double(), sd[key + ".running_var"].double()
    s = g / torch.sqrt(var + eps)
    return s, b - mu * s


class SwinUPerNetEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], enc_prefix: str, dec_prefix: str, cfg: SwinCfg,
                 device: torch.device, max_batch: int = 8, norm_mean: Optional[Sequence[float]] = None,
                 norm_std: Optional[Sequence[float]] = None):
        if device.type != "cuda":
            raise nv.NativeError("SwinUPerNetEngine needs a CUDA device (no CPU fallback)")
        if cfg.embed_dim % 128 != 0:
            raise NotImplementedError(f"embed_dim {cfg.embed_dim}: the sm_100a kernels are tiled for multiples of 128")
        if cfg.patch % 128 != 0:
            raise NotImplementedError(f"patch {cfg.patch}: UPerNet plan needs a multiple of 128")
        nv.lib()
        self.cfg, self.dev, self.B = cfg, device, max_batch
        sd = {k: v.detach().to("cpu") for k, v in state_dict.items()}
        E, D, dev = enc_prefix, dec_prefix, device
        C0 = cfg.embed_dim
        self.dims = [C0 * 2 ** i for i in range(4)]
        self.hw = [(cfg.patch // 4) >> i for i in range(4)]

        # ---- patch embedding (same packing as the ConvNeXt stem: [64][C0], k = ky*16 + kx*4 + c)
        w = sd[E + "patch_embed.proj.weight"].float()
        assert w.shape == (C0, cfg.in_chans, 4, 4), w.shape
        wk = torch.zeros(4, 4, 4, C0, dtype=torch.float64)
        wk[:, :, :cfg.in_chans, :] = w.double().permute(2, 3, 1, 0)
        self.stem_w_f32 = _f32(wk.reshape(64, C0), dev)
        self.stem_b_f32 = _f32(sd[E + "patch_embed.proj.bias"], dev)
        if norm_mean is not None:
            mean = torch.zeros(4, dtype=torch.float64)
            std = torch.ones(4, dtype=torch.float64)
            mean[:cfg.in_chans] = torch.tensor(list(norm_mean), dtype=torch.float64)
            std[:cfg.in_chans] = torch.tensor(list(norm_std), dtype=torch.float64)
            wf = wk / std.view(1, 1, 4, 1)
            bf = sd[E + "patch_embed.proj.bias"].double() - (wf * mean.view(1, 1, 4, 1)).sum(dim=(0, 1, 2))
            self.stem_w_u8, self.stem_b_u8 = _f32(wf.reshape(64, C0), dev), _f32(bf, dev)
        self.stem_ln_w = _f32(sd[E + "patch_embed.norm.weight"], dev)
        self.stem_ln_b = _f32(sd[E + "patch_embed.norm.bias"], dev)

        # ---- stages
        self.stages: List[dict] = []
        for i, (depth, C, nh, res) in enumerate(zip(cfg.depths, self.dims, cfg.heads, self.hw)):
            S = E + f"layers.{i}."
            # timm _calc_window_shift: window shrinks to the resolution (and the shift vanishes) when res <= window
            w_eff = res if res <= cfg.window else cfg.window
            s_eff = 0 if res <= cfg.window else cfg.window // 2
            if w_eff > 12 or C != nh * 32:
                raise NotImplementedError(f"stage {i}: window {w_eff} / head dim {C // nh} outside the attention kernel")
            st = {"C": C, "heads": nh, "window": w_eff, "blocks": []}
            if i > 0:
                st["mg_ln_w"] = _f32(sd[S + "downsample.norm.weight"], dev)
                st["mg_ln_b"] = _f32(sd[S + "downsample.norm.bias"], dev)
                st["mg_w"] = _bf16(sd[S + "downsample.reduction.weight"], dev)          # [C, 4*C_prev], no bias
                st["mg_b"] = torch.zeros(C, dtype=torch.float32, device=dev)
            for j in range(depth):
                Bk = S + f"blocks.{j}."
                table = sd[Bk + "attn.relative_position_bias_table"].float()            # [(2w-1)^2, heads]
                assert table.shape == ((2 * w_eff - 1) ** 2, nh), (table.shape, w_eff, nh)
                qkv_b = sd[Bk + "attn.qkv.bias"].float()
                st["blocks"].append({
                    "shift": 0 if j % 2 == 0 else s_eff,
                    "n1_w": _f32(sd[Bk + "norm1.weight"], dev), "n1_b": _f32(sd[Bk + "norm1.bias"], dev),
                    "qkv_w": _bf16(sd[Bk + "attn.qkv.weight"], dev), "qkv_b": _f32(qkv_b, dev),
                    "qkv_b_bf16": _bf16(qkv_b, dev),                                     # q/k/v of a padded token
                    "table": _f32(table.t(), dev),                                       # [heads][(2w-1)^2]
                    "proj_w": _bf16(sd[Bk + "attn.proj.weight"], dev), "proj_b": _f32(sd[Bk + "attn.proj.bias"], dev),
                    "n2_w": _f32(sd[Bk + "norm2.weight"], dev), "n2_b": _f32(sd[Bk + "norm2.bias"], dev),
                    "fc1_w": _bf16(sd[Bk + "mlp.fc1.weight"], dev), "fc1_b": _f32(sd[Bk + "mlp.fc1.bias"], dev),
                    "fc2_w": _bf16(sd[Bk + "mlp.fc2.weight"], dev), "fc2_b": _f32(sd[Bk + "mlp.fc2.bias"], dev),
                })
            self.stages.append(st)
        self.scale = 32 ** -0.5

        # ---- UPerNet decoder
        Pc, Sc = cfg.pyramid_channels, cfg.segmentation_channels
        assert Pc % 64 == 0 and Sc % 64 == 0

        def conv1x1_bn(prefix):          # smp Conv2dReLU(k=1): conv (no bias) + BN + ReLU -> GEMM weights
            wt = sd[prefix + ".0.weight"].double()
            s, b = _bn_fold(sd, prefix + ".1")
            return _bf16((wt[:, :, 0, 0] * s.view(-1, 1)).float(), dev), _f32(b, dev)

        Dd = D + "decoder."
        c32 = self.dims[3]
        self.psp = [conv1x1_bn(Dd + f"psp.blocks.{k}.1") for k in range(len(PSP_SIZES))]
        self.psp_out = conv1x1_bn(Dd + "psp.out_conv")
        assert self.psp_out[0].shape == (Pc, 2 * c32), self.psp_out[0].shape
        # fpn_stages 0..2 take f16, f8, f4; stage 3 is the 0-channel dummy (Identity), stage 4 is never reached
        self.lateral = [conv1x1_bn(Dd + f"fpn_stages.{k}.skip_conv") for k in range(3)]
        wf = sd[Dd + "fpn_bottleneck.0.weight"].float()                                  # [Sc, 5*Pc, 3, 3]
        assert wf.shape == (Sc, 5 * Pc, 3, 3), wf.shape
        s, b = _bn_fold(sd, Dd + "fpn_bottleneck.1")
        self.fuse_w, self.fuse_s, self.fuse_b = _bf16(wf.permute(0, 2, 3, 1), dev), _f32(s, dev), _f32(b, dev)
        wh = sd[D + "segmentation_head.0.weight"].float()                                # [n_cls, Sc, 1, 1]
        assert wh.shape[:2] == (cfg.n_classes, Sc) and cfg.n_classes <= 64
        whp = torch.zeros(64, Sc)
        whp[:cfg.n_classes] = wh[:, :, 0, 0]
        bh = torch.zeros(64)
        bh[:cfg.n_classes] = sd[D + "segmentation_head.0.bias"].float()
        self.head_w, self.head_b = _bf16(whp, dev), _f32(bh, dev)
        self._alloc_workspace()

    # ------------------------------------------------------------------------------ workspace
    def _alloc_workspace(self):
        cfg, B, dev = self.cfg, self.B, self.dev
        bf, f32 = nv.op_dtype(), torch.float32
        hw, dims = self.hw, self.dims
        Pc, Sc = cfg.pyramid_channels, cfg.segmentation_channels
        self.x = [torch.empty((B, h, h, c), dtype=f32, device=dev) for h, c in zip(hw, dims)]
        tc_max = max(h * h * c for h, c in zip(hw, dims))
        self.y = torch.empty(B * tc_max, dtype=bf, device=dev)             # LN output / merge operand / attention out
        self.a = torch.empty(B * tc_max, dtype=bf, device=dev)             # attention output
        self.qkv = torch.empty(B * tc_max * 3, dtype=bf, device=dev)
        self.h = torch.empty(B * tc_max * 4, dtype=bf, device=dev)         # MLP hidden
        self.fb = [torch.empty((B, h, h, c), dtype=bf, device=dev) for h, c in zip(hw, dims)]   # decoder operands
        self.pool = [torch.empty((B, s, s, dims[3]), dtype=bf, device=dev) for s in PSP_SIZES]
        self.pool_c = [torch.empty((B, s, s, dims[3] // 4), dtype=bf, device=dev) for s in PSP_SIZES]
        self.psp_cat = torch.empty((B, hw[3], hw[3], 2 * dims[3]), dtype=bf, device=dev)
        self.p = [torch.empty((B, h, h, Pc), dtype=bf, device=dev) for h in hw[::-1]]            # P0 (deepest) .. P3
        self.lat = torch.empty(B * hw[0] * hw[0] * Pc, dtype=bf, device=dev)
        self.cat = torch.empty((B, hw[0], hw[0], 5 * Pc), dtype=bf, device=dev)
        self.fused = torch.empty((B, hw[0], hw[0], Sc), dtype=bf, device=dev)
        self.logits_q = torch.empty((B, hw[0], hw[0], 64), dtype=f32, device=dev)                # at H/4, 64-wide rows
        self.logits = torch.empty((B, cfg.n_classes, cfg.patch, cfg.patch), dtype=f32, device=dev)

    # ------------------------------------------------------------------------------ encoder
    def _encode(self, n: int) -> None:
        for i, st in enumerate(self.stages):
            C, hwi, nh, ws = st["C"], self.hw[i], st["heads"], st["window"]
            T = n * hwi * hwi
            x = self.x[i][:n]
            xm = x.view(T, C)
            if i > 0:
                Cp = self.dims[i - 1]
                mg = self.y[:T * 4 * Cp].view(n, hwi, hwi, 4 * Cp)
                nv.merge_ln(self.x[i - 1][:n], st["mg_ln_w"], st["mg_ln_b"], mg)
                nv.gemm_bf16(mg.view(T, 4 * Cp), st["mg_w"], nv.EPI_F32, bias=st["mg_b"], out=xm)
            y = self.y[:T * C].view(T, C)
            qkv = self.qkv[:T * 3 * C].view(n, hwi, hwi, 3 * C)
            a = self.a[:T * C].view(n, hwi, hwi, C)
            hbuf = self.h[:T * 4 * C].view(T, 4 * C)
            R = nv.EPI_REVERSE_TILES
            for blk in st["blocks"]:
                nv.layernorm_rows(xm, blk["n1_w"], blk["n1_b"], y)
                # GEMMs that consume a tensor the previous kernel has just streamed out walk their tiles backwards
                # (newest rows are still in L2) and leave the rows the next forward kernel starts with for last
                nv.gemm_bf16(y, blk["qkv_w"], nv.EPI_BF16 | R, bias=blk["qkv_b"], out=qkv.view(T, 3 * C))
                nv.swin_window_attn(qkv, blk["qkv_b_bf16"], blk["table"], a, nh, ws, blk["shift"], self.scale)
                nv.gemm_bf16(a.view(T, C), blk["proj_w"], nv.EPI_RESID_F32 | R, bias=blk["proj_b"], resid=xm, out=xm)
                nv.layernorm_rows(xm, blk["n2_w"], blk["n2_b"], y)
                nv.gemm_bf16(y, blk["fc1_w"], nv.EPI_GELU_BF16 | R, bias=blk["fc1_b"], out=hbuf)
                nv.gemm_bf16(hbuf, blk["fc2_w"], nv.EPI_RESID_F32, bias=blk["fc2_b"], resid=xm, out=xm)

    def encode_u8(self, tiles_u8: torch.Tensor) -> None:
        if not hasattr(self, "stem_w_u8"):
            raise nv.NativeError("engine was built without normalisation constants: uint8 input unavailable")
        n = tiles_u8.shape[0]
        assert n <= self.B
        nv.stem_ln(tiles_u8, self.stem_w_u8, self.stem_b_u8, self.stem_ln_w, self.stem_ln_b, self.x[0][:n], eps=1e-5)
        self._encode(n)

    def encode_f32(self, x_nchw: torch.Tensor) -> None:
        n = x_nchw.shape[0]
        assert n <= self.B and x_nchw.shape[1] == self.cfg.in_chans
        nv.stem_ln_f32(x_nchw, self.stem_w_f32, self.stem_b_f32, self.stem_ln_w, self.stem_ln_b, self.x[0][:n],
                       eps=1e-5)
        self._encode(n)

    def features(self, n: int) -> List[torch.Tensor]:
        """Stage outputs (strides 4/8/16/32) as fp32 NHWC views (no final norm, like timm's feature getter)."""
        return [x[:n] for x in self.x]

    # ------------------------------------------------------------------------------ decoder
    def _decode(self, n: int, out: Optional[torch.Tensor] = None, quarter: bool = False) -> torch.Tensor:
        """-> fp32 logits [n, n_cls, P, P] (in ``out`` when given and contiguous, else in the engine's buffer); with
        ``quarter`` the head's x4 upsampling is left to the consumer: returns [n, P/4, P/4, 64] (first n_cls valid)."""
        cfg, hw, dims = self.cfg, self.hw, self.dims
        Pc = cfg.pyramid_channels
        fb = [t[:n] for t in self.fb]
        for i in range(4):
            nv.cast_f32_bf16(self.x[i][:n], fb[i])
        # PSP on the stride-32 map
        cat = self.psp_cat[:n]
        nv.bilinear_slice(fb[3], cat, 0)
        for k, s in enumerate(PSP_SIZES):
            pooled, pc = self.pool[k][:n], self.pool_c[k][:n]
            nv.adaptive_avgpool(fb[3], s, pooled)
            nv.gemm_bf16(pooled.view(n * s * s, dims[3]), self.psp[k][0], nv.EPI_RELU_BF16, bias=self.psp[k][1],
                         out=pc.view(n * s * s, dims[3] // 4))
            nv.bilinear_slice(pc, cat, dims[3] + k * (dims[3] // 4))
        p = [t[:n] for t in self.p]
        nv.gemm_bf16(cat.view(n * hw[3] * hw[3], 2 * dims[3]), self.psp_out[0], nv.EPI_RELU_BF16, bias=self.psp_out[1],
                     out=p[0].view(-1, Pc))
        # FPN top-down: P_{k+1} = bilinear(P_k -> skip size) + relu(bn(conv1x1(skip)))
        for k in range(3):
            f = fb[2 - k]
            h = hw[2 - k]
            lat = self.lat[:n * h * h * Pc].view(n, h, h, Pc)
            nv.gemm_bf16(f.view(n * h * h, dims[2 - k]), self.lateral[k][0], nv.EPI_RELU_BF16, bias=self.lateral[k][1],
                         out=lat.view(-1, Pc))
            nv.bilinear_slice(p[k], p[k + 1], 0, add=lat)
        # five maps at H/4: P0..P3 resized, and down2(up2(P3)) for the 0-channel stage
        big = self.cat[:n]
        if Pc % 8 == 0 and 256 % (Pc // 8) == 0:
            nv.pyramid_concat(p[0], p[1], p[2], p[3], big)          # one contiguous write of all five slices
        else:
            for k in range(4):
                nv.bilinear_slice(p[k], big, k * Pc)
            nv.updown_slice(p[3], big, 4 * Pc)
        fused = self.fused[:n]
        nv.conv3x3(big, self.fuse_w, self.fuse_s, self.fuse_b, nv.CONV_RELU_BF16, out=fused)
        lq = self.logits_q[:n]
        nv.gemm_bf16(fused.view(-1, cfg.segmentation_channels), self.head_w, nv.EPI_F32, bias=self.head_b,
                     out=lq.view(-1, 64))
        if quarter:
            return lq
        direct = out is not None and out.is_contiguous() and out.dtype == torch.float32
        dst = out if direct else self.logits[:n]
        nv.head_upsample4(lq, cfg.n_classes, dst)
        if out is not None and not direct:
            out.copy_(dst)
            return out
        return dst

    def decode_logits_nchw(self, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 logits [n, n_classes, P, P] -- the layout FLAIR_HUB_Model.forward returns."""
        if out is None:
            out = torch.empty((n, self.cfg.n_classes, self.cfg.patch, self.cfg.patch), dtype=torch.float32,
                              device=self.dev)
        return self._decode(n, out)

    def decode_argmax_to_raster(self, n: int, plan: torch.Tensor, own: Optional[torch.Tensor], raster: torch.Tensor,
                                margin: int) -> None:
        """inference.py:295-352 on the interpolated logits (argmax must follow the x4 bilinear)."""
        lq = self._decode(n, quarter=True)     # the full-resolution logits (737 MB per 37 tiles) are never stored
        nv.crop_argmax_write(lq, nv.NHWC_UP4, margin, plan[:n], own[:n] if own is not None else None, raster,
                             n_cls=self.cfg.n_classes)

    def launch_count(self) -> int:
        """Kernel launches of one batch after the tile gather (bench.py's gpu_launches)."""
        n = 1                                                   # patch embed
        for i, d in enumerate(self.cfg.depths):
            n += (2 if i > 0 else 0) + 7 * d
        resizes = 1 if 256 % max(self.cfg.pyramid_channels // 8, 1) == 0 else 5        # fused pyramid concat
        n += 4 + 1 + 3 * len(PSP_SIZES) + 1 + 2 * 3 + resizes + 1 + 1 + 1   # casts, PSP, FPN, resizes, fuse, head, crop(x4)
        return n

"""Oracle (test infrastructure): fp32 eager restatement of timm ``swin_base_patch4_window12_384`` (as
smp's ``TimmUniversalEncoder`` runs it at ``img_size=512``) and of smp 0.4.0's ``UPerNetDecoder`` +
``SegmentationHead(kernel_size=1, upsampling=4)`` -- BASELINE.json configs[2], SURVEY.md section 8 A5 and
appendix D.  Call sites in the reference: flair_hub/models/monotemp_model.py:68-92 (``smp.create_model``),
flair_hub/models/flair_model.py:376 (encoder forward), :418 (decoder forward).

Both packages are absent from this image; PARITY UNPINNED by the reference (it has no tests).  What is
pinned here (tests/test_oracle_swin.py): the Swin encoder against HuggingFace ``SwinModel`` for every
shape on which the two libraries agree by construction (no padding, or padding in un-shifted blocks only).

Padding semantics chosen: recent timm -- ``norm1`` -> cyclic shift (``torch.roll``)
-> zero-pad bottom/right to a multiple of the window -> partition -> attention (padded tokens are ordinary
tokens: q/k/v = the qkv bias; they are masked only by the shift mask, which is built on the PADDED grid with
the usual (0,-ws), (-ws,-shift), (-shift,None) slices) -> window reverse -> crop -> roll back.  Patch merging
is at the START of stages 1..3 (timm >= 0.9), concat order (h0w0, h1w0, h0w1, h1w1), LayerNorm(4C) then a
bias-free Linear.  ``nn.LayerNorm`` eps = 1e-5 throughout.  No final norm on the returned features.

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may import this file.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

SWIN_CFGS = {
    # name: (embed_dim, depths, heads, window)
    "swin_tiny_patch4_window7_224": (96, (2, 2, 6, 2), (3, 6, 12, 24), 7),
    "swin_small_patch4_window7_224": (96, (2, 2, 18, 2), (3, 6, 12, 24), 7),
    "swin_base_patch4_window7_224": (128, (2, 2, 18, 2), (4, 8, 16, 32), 7),
    "swin_base_patch4_window12_384": (128, (2, 2, 18, 2), (4, 8, 16, 32), 12),
}


def window_partition(x: torch.Tensor, ws: int) -> torch.Tensor:
    b, h, w, c = x.shape
    x = x.view(b, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, c)


def window_reverse(windows: torch.Tensor, ws: int, h: int, w: int) -> torch.Tensor:
    c = windows.shape[-1]
    x = windows.view(-1, h // ws, w // ws, ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, h, w, c)


def relative_position_index(ws: int) -> torch.Tensor:
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)  # 2, N
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)  # N, N


class WindowAttention(nn.Module):
    def __init__(self, dim: int, heads: int, ws: int):
        super().__init__()
        self.heads, self.ws = heads, ws
        self.scale = (dim // heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, heads))
        self.register_buffer("relative_position_index", relative_position_index(ws), persistent=False)
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def rel_pos_bias(self) -> torch.Tensor:
        n = self.ws * self.ws
        b = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(n, n, -1)
        return b.permute(2, 0, 1).contiguous().unsqueeze(0)  # 1, heads, N, N

    def forward(self, x, mask=None):
        b_, n, c = x.shape
        qkv = self.qkv(x).reshape(b_, n, 3, self.heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = attn + self.rel_pos_bias()
        if mask is not None:
            nw = mask.shape[0]
            attn = attn.view(-1, nw, self.heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
            attn = attn.view(-1, self.heads, n, n)
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(b_, n, -1)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


def shift_attn_mask(h: int, w: int, ws: int, shift: int) -> torch.Tensor:
    """(num_windows, N, N) additive mask (0 / -100) on the grid padded up to a multiple of ws."""
    hp, wp = math.ceil(h / ws) * ws, math.ceil(w / ws) * ws
    img = torch.zeros((1, hp, wp, 1))
    cnt = 0
    for hs in ((0, -ws), (-ws, -shift), (-shift, None)):
        for wsl in ((0, -ws), (-ws, -shift), (-shift, None)):
            img[:, hs[0]:hs[1], wsl[0]:wsl[1], :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


class SwinBlock(nn.Module):
    def __init__(self, dim: int, heads: int, ws: int, shift: int):
        super().__init__()
        self.ws, self.shift = ws, shift
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, heads, ws)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, dim * 4)

    def _attn(self, x):
        b, h, w, c = x.shape
        ws, shift = self.ws, self.shift
        sx = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2)) if shift else x
        pad_h, pad_w = (ws - h % ws) % ws, (ws - w % ws) % ws
        sx = F.pad(sx, (0, 0, 0, pad_w, 0, pad_h))
        hp, wp = h + pad_h, w + pad_w
        xw = window_partition(sx, ws).view(-1, ws * ws, c)
        mask = shift_attn_mask(h, w, ws, shift).to(x) if shift else None
        aw = self.attn(xw, mask).view(-1, ws, ws, c)
        sx = window_reverse(aw, ws, hp, wp)[:, :h, :w, :].contiguous()
        return torch.roll(sx, shifts=(shift, shift), dims=(1, 2)) if shift else sx

    def forward(self, x):  # B,H,W,C
        x = x + self._attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchMerging(nn.Module):
    def __init__(self, dim: int, out_dim: int):
        super().__init__()
        self.norm = nn.LayerNorm(4 * dim)
        self.reduction = nn.Linear(4 * dim, out_dim, bias=False)

    def forward(self, x):
        b, h, w, c = x.shape
        x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2))
        _, h, w, _ = x.shape
        x = x.reshape(b, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 4, 2, 5).flatten(3)
        return self.reduction(self.norm(x))


class SwinStage(nn.Module):
    def __init__(self, dim: int, out_dim: int, res: int, depth: int, heads: int, ws: int, downsample: bool):
        super().__init__()
        self.downsample = PatchMerging(dim, out_dim) if downsample else nn.Identity()
        # timm _calc_window_shift: the window shrinks to the resolution (and the shift vanishes) when res <= window
        w_eff = res if res <= ws else ws
        s_eff = 0 if res <= ws else ws // 2
        self.blocks = nn.Sequential(*[SwinBlock(out_dim, heads, w_eff, 0 if i % 2 == 0 else s_eff)
                                      for i in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class PatchEmbed(nn.Module):
    def __init__(self, in_chans: int, dim: int):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=4, stride=4)
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(self.proj(x).permute(0, 2, 3, 1))


class SwinTransformer(nn.Module):
    """timm SwinTransformer pruned by FeatureGetterNet (no final norm, no head); NHWC internally."""

    def __init__(self, in_chans: int, img_size: int, embed_dim: int, depths: Sequence[int], heads: Sequence[int],
                 window: int):
        super().__init__()
        self.patch_embed = PatchEmbed(in_chans, embed_dim)
        layers = []
        dim, res = embed_dim, img_size // 4
        for i, (d, nh) in enumerate(zip(depths, heads)):
            out_dim = embed_dim * 2 ** i
            if i > 0:
                res //= 2
            layers.append(SwinStage(dim, out_dim, res, d, nh, window, downsample=i > 0))
            dim = out_dim
        self.layers = nn.Sequential(*layers)

    def forward_intermediates(self, x) -> List[torch.Tensor]:
        x = self.patch_embed(x)
        feats = []
        for stage in self.layers:
            x = stage(x)
            feats.append(x.permute(0, 3, 1, 2).contiguous())
        return feats


class _FeatureGetter(nn.Module):
    """timm ``FeatureGetterNet``: keeps the backbone as ``.model`` (hence the ``model.model.`` key prefix)."""

    def __init__(self, model: SwinTransformer):
        super().__init__()
        self.model = model

    def forward(self, x):
        return self.model.forward_intermediates(x)


class SwinUniversalEncoder(nn.Module):
    """smp ``TimmUniversalEncoder`` around a Swin backbone: [x, empty(B,0,H/2,W/2), f4, f8, f16, f32]."""

    def __init__(self, name: str, in_channels: int, img_size: int = 512):
        super().__init__()
        dim, depths, heads, window = SWIN_CFGS[name]
        self.model = _FeatureGetter(SwinTransformer(in_channels, img_size, dim, depths, heads, window))
        self.out_channels = [in_channels, 0] + [dim * 2 ** i for i in range(4)]
        self.output_stride = 32

    def forward(self, x):
        feats = self.model(x)
        b, _, h, w = x.shape
        return [x, torch.empty([b, 0, h // 2, w // 2], dtype=x.dtype, device=x.device)] + feats


# --------------------------------------------------------------------------------------
# smp 0.4.0 UPerNet decoder
# --------------------------------------------------------------------------------------
def conv_bn_relu(cin: int, cout: int, k: int) -> nn.Sequential:
    """smp ``Conv2dReLU(use_batchnorm=True)``: conv k x k (no bias, pad k//2) + BatchNorm2d + ReLU."""
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class PSPModule(nn.Module):
    def __init__(self, cin: int, cout: int, sizes=(1, 2, 3, 6)):
        super().__init__()
        self.blocks = nn.ModuleList([nn.Sequential(nn.AdaptiveAvgPool2d(s), conv_bn_relu(cin, cin // len(sizes), 1))
                                     for s in sizes])
        self.out_conv = conv_bn_relu(cin * 2, cout, 1)

    def forward(self, x):
        h, w = x.shape[2:]
        out = [x] + [F.interpolate(blk(x), size=(h, w), mode="bilinear", align_corners=False) for blk in self.blocks]
        return self.out_conv(torch.cat(out, dim=1))


class FPNBlock(nn.Module):
    def __init__(self, skip_channels: int, pyramid_channels: int):
        super().__init__()
        self.skip_conv = conv_bn_relu(skip_channels, pyramid_channels, 1) if skip_channels != 0 else nn.Identity()

    def forward(self, x, skip):
        _, ch, h, w = skip.shape
        x = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)
        if ch != 0:
            x = x + self.skip_conv(skip)
        return x


class UPerNetDecoder(nn.Module):
    def __init__(self, encoder_channels: Sequence[int], pyramid_channels: int = 256, segmentation_channels: int = 64):
        super().__init__()
        enc = list(encoder_channels)[::-1]          # [c32, c16, c8, c4, 0, C_in]
        self.psp = PSPModule(enc[0], pyramid_channels)
        # smp builds one FPN block per remaining entry, including one for the input-resolution feature (C_in
        # channels) that forward() never reaches (zip stops after the 0-channel dummy): its parameters exist in
        # checkpoints but take no part in the arithmetic.
        self.fpn_stages = nn.ModuleList([FPNBlock(ch, pyramid_channels) for ch in enc[1:]])
        self.fpn_bottleneck = conv_bn_relu((len(enc) - 1) * pyramid_channels, segmentation_channels, 3)

    def forward(self, *features):
        out_h, out_w = features[0].shape[2:]
        target = (out_h // 4, out_w // 4)
        feats = features[1:][::-1]
        fpn = [self.psp(feats[0])]
        for f, stage in zip(feats[1:], self.fpn_stages):
            fpn.append(stage(fpn[-1], f))
        resized = [F.interpolate(f, size=target, mode="bilinear", align_corners=False) for f in fpn]
        return self.fpn_bottleneck(torch.cat(resized, dim=1))


class UPerNetHead(nn.Sequential):
    """smp ``SegmentationHead(in, classes, kernel_size=1, upsampling=4)``: conv1x1 (bias) then
    ``nn.UpsamplingBilinear2d(scale_factor=4)`` (= bilinear, align_corners=True)."""

    def __init__(self, cin: int, classes: int):
        super().__init__(nn.Conv2d(cin, classes, 1), nn.UpsamplingBilinear2d(scale_factor=4))

"""smp 0.4.0 ``UnetDecoder`` + ``SegmentationHead`` execution plan, shared by every encoder family.

Replaces ``main_decoders[task].seg_model(*features)`` at flair_hub/models/flair_model.py:417-419
(monotemp_model.py:22-31).  Block k: nearest x2 upsample of the running tensor, concat with the skip (when it
has channels), conv3x3+BN+ReLU twice; head conv3x3 with one of three epilogues (fp32 logits NCHW / NHWC, or
crop + argmax + ownership write straight into the zone raster).  Eval BatchNorm is an fp32 scale/bias in the conv
epilogue, so the bf16 weights are exactly the checkpoint's.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from .. import native as nv


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


class UNetDecoderPlan:
    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, encoder_channels: Sequence[int], n_classes: int,
                 patch: int, max_batch: int, device: torch.device, decoder_channels: Sequence[int] = (256, 128, 64, 32, 16),
                 bn_eps: float = 1e-5):
        self.dev, self.B, self.P, self.n_classes = device, max_batch, patch, n_classes
        enc = list(encoder_channels)[1:][::-1]            # smp: drop the stride-1 feature, deepest first
        in_ch = [enc[0]] + list(decoder_channels[:-1])
        skip_ch = list(enc[1:]) + [0]
        self.blocks: List[dict] = []
        for k, (ci, cs, co) in enumerate(zip(in_ch, skip_ch, decoder_channels)):
            blk = {"cin": ci, "cskip": cs, "cout": co}
            for name in ("conv1", "conv2"):
                kp = prefix + f"decoder.blocks.{k}.{name}."
                w = sd[kp + "0.weight"].float()                     # [co, cin_total, 3, 3]
                g, b_ = sd[kp + "1.weight"].double(), sd[kp + "1.bias"].double()
                mu, var = sd[kp + "1.running_mean"].double(), sd[kp + "1.running_var"].double()
                scale = g / torch.sqrt(var + bn_eps)
                blk[name + "_w"] = _bf16(w.permute(0, 2, 3, 1), device)   # [co][3][3][cin]
                blk[name + "_s"] = _f32(scale, device)
                blk[name + "_b"] = _f32(b_ - mu * scale, device)
            assert blk["conv1_w"].shape[-1] == ci + cs, (blk["conv1_w"].shape, ci, cs)
            if (cs > 0 and ci % 64 == 0 and cs % 64 == 0 and co % 64 == 0
                    and os.environ.get("FZ_CATCONV", "1") != "0"):
                # skip block: conv3x3(cat(up2(a), skip)) as one implicit GEMM (csrc/catconv3x3_tcgen05.cu)
                w1 = sd[prefix + f"decoder.blocks.{k}.conv1.0.weight"]
                blk["conv1_w16a"] = _bf16(nv.merge_upconv_weights(w1[:, :ci]), device)
            if cs == 0 and ci in (32, 64) and co in (16, 32) and os.environ.get("FZ_UPCONV", "1") != "0":
                # no skip: conv3x3(nearest_up2(a)) straight from `a` with merged sub-pixel taps (csrc/upconv3x3_rows.cu)
                w1 = sd[prefix + f"decoder.blocks.{k}.conv1.0.weight"]
                blk["conv1_w16"] = _bf16(nv.merge_upconv_weights(w1), device)
            self.blocks.append(blk)
        wh = sd[prefix + "segmentation_head.0.weight"].float()      # [ncls, c_last, 3, 3]
        assert wh.shape[0] == n_classes and n_classes <= 32
        whp = torch.zeros(32, 3, 3, wh.shape[1])
        whp[:n_classes] = wh.permute(0, 2, 3, 1)
        bh = torch.zeros(32)
        bh[:n_classes] = sd[prefix + "segmentation_head.0.bias"].float()
        self.head_w, self.head_b = _bf16(whp, device), _f32(bh, device)
        # workspace: concat buffer + two ping-pong conv outputs
        deepest = patch >> (len(decoder_channels))                  # stride-32 map
        hd, cat_sz, out_sz = deepest, [], []
        for blk in self.blocks:
            hd *= 2
            cat_sz.append(hd * hd * (blk["cin"] + blk["cskip"]))
            out_sz.append(hd * hd * blk["cout"])
        self.deepest = deepest
        bf = nv.op_dtype()
        self.cat = torch.empty(max_batch * max(cat_sz), dtype=bf, device=device)
        self.t1 = torch.empty(max_batch * max(out_sz), dtype=bf, device=device)
        self.t2 = torch.empty(max_batch * max(out_sz), dtype=bf, device=device)

    def _body(self, feats_deep_first: List[torch.Tensor], n: int) -> torch.Tensor:
        """feats_deep_first: NHWC feature maps, deepest (stride 32) first, one per skip that has channels."""
        a = feats_deep_first[0]
        skips = list(feats_deep_first[1:])
        hd = self.deepest
        for k, blk in enumerate(self.blocks):
            hd *= 2
            ct = blk["cin"] + blk["cskip"]
            o1 = self.t1[:n * hd * hd * blk["cout"]].view(n, hd, hd, blk["cout"])
            hs = hd // 2
            skip_k = skips[k] if blk["cskip"] > 0 else None
            if ("conv1_w16" in blk and a.dtype == nv.op_dtype() and hs % 128 == 0 and a.is_contiguous()):
                nv.upconv3x3_bn_relu(a, blk["conv1_w16"], blk["conv1_s"], blk["conv1_b"], o1)
            elif ("conv1_w16a" in blk and a.dtype == nv.op_dtype() and skip_k is not None
                  and skip_k.dtype == nv.op_dtype() and hs * hs >= 128 and (hs >= 128 and hs % 128 == 0 or 128 % hs == 0)
                  and a.is_contiguous() and skip_k.is_contiguous()):
                nv.catconv3x3_bn_relu(a, skip_k, blk["conv1_w16a"], blk["conv1_w"], blk["conv1_s"], blk["conv1_b"], o1)
            else:
                cat = self.cat[:n * hd * hd * ct].view(n, hd, hd, ct)
                skip = skips[k] if blk["cskip"] > 0 else None
                nv.upsample2_concat(a, skip, cat)
                nv.conv3x3(cat, blk["conv1_w"], blk["conv1_s"], blk["conv1_b"], nv.CONV_RELU_BF16, out=o1)
            o2 = self.t2[:n * hd * hd * blk["cout"]].view(n, hd, hd, blk["cout"])
            nv.conv3x3(o1, blk["conv2_w"], blk["conv2_s"], blk["conv2_b"], nv.CONV_RELU_BF16, out=o2)
            a = o2
        return a

    def logits_nchw(self, feats, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        a = self._body(feats, n)
        if out is None:
            out = torch.empty((n, self.n_classes, self.P, self.P), dtype=torch.float32, device=self.dev)
        nv.conv3x3(a, self.head_w, None, self.head_b, nv.CONV_LOGITS_F32_NCHW, out=out, cout=self.n_classes)
        return out

    def logits_nhwc(self, feats, n: int, out: torch.Tensor) -> torch.Tensor:
        a = self._body(feats, n)
        nv.conv3x3(a, self.head_w, None, self.head_b, nv.CONV_LOGITS_F32, out=out, cout=self.n_classes,
                   cstride=out.shape[-1])
        return out

    def argmax_to_raster(self, feats, n: int, plan, own, raster, margin: int) -> None:
        a = self._body(feats, n)
        nv.conv3x3(a, self.head_w, None, self.head_b, nv.CONV_ARGMAX_RASTER, cout=self.n_classes, plan=plan, own=own,
                   raster=raster, margin=margin)

    def launches(self) -> int:
        hd, n = self.deepest, 1
        for blk in self.blocks:
            hd *= 2
            hs = hd // 2
            fused = ("conv1_w16" in blk and hs % 128 == 0) or \
                    ("conv1_w16a" in blk and hs * hs >= 128 and (hs % 128 == 0 or 128 % hs == 0))
            n += 2 if fused else 3        # (up-)conv1 [+ upsample/concat], conv2 -- with bf16 feature maps
        return n

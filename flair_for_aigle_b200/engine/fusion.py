"""Several mono-temporal encoders fused by ``FusionHandler`` (flair_hub/models/flair_model.py:473-547, the >= 2 keys
branch) in front of one U-Net decoder -- the model of BASELINE.json configs[4] (AERIAL_RGBI 4 ch + DEM_ELEV 1 ch), forward
only.

Per stage the reference concatenates the modalities' feature maps along channels and applies ``conv_f[i]`` (1x1, bias):
    fused_i = W_i [x_i^(1) ; x_i^(2) ; ...] + b_i = sum_m W_i[:, slice_m] x_i^(m) + b_i
so no concatenated tensor is materialised: one tcgen05 GEMM per (stage, modality), the first with the bias epilogue
(FZ_EPI_F32), the following ones accumulating through the fp32 residual epilogue (FZ_EPI_RESID_F32, in place).
A modality whose maps are smaller / larger than the first one's (another patch size, e.g. a coarser DEM) is aligned with
the bilinear resize of flair_model.py:523-529 on its bf16 operand.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .. import native as nv
from .unet_decoder import _bf16, _f32


class FusedEncodersUNet:
    """encoders: {modality: engine} in FusionHandler order (= mono_keys order); every engine exposes ``encode_f32``,
    ``features(n)`` (fp32 NHWC, strides 4/8/16/32), ``decoder`` (UNetDecoderPlan), ``B``, ``cfg``, ``dev``."""

    def __init__(self, encoders: Dict[str, object], state_dict: Dict[str, torch.Tensor], fusion_prefix: str):
        self.mods = list(encoders.keys())
        self.enc = encoders
        first = encoders[self.mods[0]]
        self.B, self.cfg, self.dev = first.B, first.cfg, first.dev
        self.decoder = first.decoder
        dev = self.dev
        feats0 = first.features(1)
        self.stage_shapes = [tuple(f.shape[1:]) for f in feats0]                 # (h, w, C_target)
        chans = {m: [f.shape[-1] for f in encoders[m].features(1)] for m in self.mods}
        # modalities at another patch size (e.g. a coarser DEM): their maps are resized to the first modality's
        # (F.interpolate bilinear, align_corners=False; flair_model.py:523-529) on the bf16 operand
        self.src_hw = {m: [tuple(f.shape[1:3]) for f in encoders[m].features(1)] for m in self.mods}
        self.w: List[List[torch.Tensor]] = []
        self.b: List[torch.Tensor] = []
        for i, (_, _, ct) in enumerate(self.stage_shapes):
            wt = state_dict[f"{fusion_prefix}conv_f.{i}.weight"].detach().to("cpu").float()
            tot = sum(chans[m][i] for m in self.mods)
            assert wt.shape[:2] == (ct, tot), (wt.shape, ct, tot)
            off, parts = 0, []
            for m in self.mods:
                parts.append(_bf16(wt[:, off:off + chans[m][i], 0, 0], dev))
                off += chans[m][i]
            self.w.append(parts)
            self.b.append(_f32(state_dict[f"{fusion_prefix}conv_f.{i}.bias"], dev))
        B = self.B
        self.fused = [torch.empty((B, h, w, c), dtype=torch.float32, device=dev) for h, w, c in self.stage_shapes]
        self.a_bufs = [torch.empty(B * h * w * max(chans[m][i] for m in self.mods), dtype=nv.op_dtype(), device=dev)
                       for i, (h, w, _) in enumerate(self.stage_shapes)]
        need = [max([sh[0] * sh[1] * chans[m][i] for m in self.mods
                     for sh in [self.src_hw[m][i]] if sh != (h, w)] + [0]) for i, (h, w, _) in enumerate(self.stage_shapes)]
        self.r_bufs = [torch.empty(B * n, dtype=nv.op_dtype(), device=dev) if n else None for n in need]

    def encode(self, batch: Dict[str, torch.Tensor]) -> int:
        n = None
        for m in self.mods:
            x = batch[m]
            if not x.is_cuda:
                raise nv.NativeError("inputs must be CUDA tensors (no CPU fallback)")
            n = x.shape[0] if n is None else n
            assert x.shape[0] == n
            self.enc[m].encode_f32(x.contiguous().float())
        for i, (h, w, ct) in enumerate(self.stage_shapes):
            T = n * h * w
            out = self.fused[i][:n].view(T, ct)
            for k, m in enumerate(self.mods):
                f = self.enc[m].features(n)[i]
                cm = f.shape[-1]
                a = self.a_bufs[i][:T * cm].view(T, cm)
                sh, sw = self.src_hw[m][i]
                if (sh, sw) == (h, w):
                    nv.cast_f32_bf16(f, a)
                else:
                    small = self.r_bufs[i][:n * sh * sw * cm].view(n, sh, sw, cm)
                    nv.cast_f32_bf16(f, small)
                    nv.bilinear_slice(small, a.view(n, h, w, cm), 0)
                if k == 0:
                    nv.gemm_bf16(a, self.w[i][k], nv.EPI_F32, bias=self.b[i], out=out)
                else:
                    nv.gemm_bf16(a, self.w[i][k], nv.EPI_RESID_F32, bias=self._zero(ct), resid=out, out=out)
        return n

    def _zero(self, c: int) -> torch.Tensor:
        z = getattr(self, "_zeros", None)
        if z is None or z.numel() < c:
            z = self._zeros = torch.zeros(max(c, 4096), dtype=torch.float32, device=self.dev)
        return z[:c]

    def decode_logits_nchw(self, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        feats = [f[:n] for f in self.fused][::-1]
        return self.decoder.logits_nchw(feats, n, out)

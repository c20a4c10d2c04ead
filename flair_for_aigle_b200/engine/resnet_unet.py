"""ResNet (BasicBlock) encoder + U-Net decoder execution plan -- `resnet34-unet`, BASELINE.json configs[0].

What the reference builds with ``smp.create_model('unet', encoder_name='resnet34')`` (native smp ResNetEncoder =
torchvision ResNet minus fc; monotemp_model.py:67-92) and runs at flair_model.py:376 / :417-419.
features = [x, relu(bn1(conv1 x)), layer1(maxpool .), layer2, layer3, layer4]  (strides 1/2/4/8/16/32).

conv1 (7x7/s2, C_in<=4) and the max-pool are CUDA-core kernels (csrc/resnet_ops.cu); every 3x3 conv of the
BasicBlocks runs on tcgen05 (csrc/conv3x3_tcgen05.cu: stride 1|2 via strided TMA boxes, eval BatchNorm as fp32
scale/bias, identity add + ReLU in the epilogue); the 1x1/s2 downsample is a 3x3 with only the centre tap set.
Activations bf16 NHWC.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from .. import native as nv
from .unet_decoder import UNetDecoderPlan, _bf16, _f32

RESNET_LAYERS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}


@dataclass
class ResNetCfg:
    layers: Sequence[int] = (3, 4, 6, 3)
    in_chans: int = 4
    n_classes: int = 19
    patch: int = 512
    decoder_channels: Sequence[int] = (256, 128, 64, 32, 16)


def _bn(sd, key, eps=1e-5):
    g, b = sd[key + ".weight"].double(), sd[key + ".bias"].double()
    mu, var = sd[key + ".running_mean"].double(), sd[key + ".running_var"].double()
    s = g / torch.sqrt(var + eps)
    return s, b - mu * s


class ResNetUNetEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], enc_prefix: str, dec_prefix: str, cfg: ResNetCfg,
                 device: torch.device, max_batch: int = 8, norm_mean: Optional[Sequence[float]] = None,
                 norm_std: Optional[Sequence[float]] = None):
        if device.type != "cuda":
            raise nv.NativeError("ResNetUNetEngine needs a CUDA device (no CPU fallback)")
        nv.lib()
        self.cfg, self.dev, self.B = cfg, device, max_batch
        sd = {k: v.detach().to("cpu") for k, v in state_dict.items()}
        E, dev = enc_prefix, device
        # conv1 [64, Cin, 7, 7] -> [196][64], k = (ky*7+kx)*4 + c
        w = sd[E + "conv1.weight"].double()
        assert w.shape == (64, cfg.in_chans, 7, 7), w.shape
        wk = torch.zeros(7, 7, 4, 64, dtype=torch.float64)
        wk[:, :, :cfg.in_chans, :] = w.permute(2, 3, 1, 0)
        s1, b1 = _bn(sd, E + "bn1")
        self.c1_w_f32 = _f32(wk.reshape(196, 64), dev)
        self.c1_s, self.c1_b_f32 = _f32(s1, dev), _f32(b1, dev)
        if norm_mean is not None:
            mean = torch.zeros(4, dtype=torch.float64)
            std = torch.ones(4, dtype=torch.float64)
            mean[:cfg.in_chans] = torch.tensor(list(norm_mean), dtype=torch.float64)
            std[:cfg.in_chans] = torch.tensor(list(norm_std), dtype=torch.float64)
            # NB zero padding of the NORMALISED image (conv pad 3) is not the same as padding raw zeros and
            # normalising; the reference pads after normalisation, so the uint8 path is exact only for
            # interior pixels.  The engine therefore always feeds conv1 the normalised float tensor.
        self.blocks: List[dict] = []
        inpl = 64
        for li, (nb, planes) in enumerate(zip(cfg.layers, (64, 128, 256, 512))):
            for j in range(nb):
                stride = 2 if (j == 0 and li > 0) else 1
                p = E + f"layer{li + 1}.{j}."
                sA, bA = _bn(sd, p + "bn1")
                sB, bB = _bn(sd, p + "bn2")
                blk = {"stride": stride, "cin": inpl, "cout": planes,
                       "w1": _bf16(sd[p + "conv1.weight"].float().permute(0, 2, 3, 1), dev), "s1": _f32(sA, dev),
                       "b1": _f32(bA, dev),
                       "w2": _bf16(sd[p + "conv2.weight"].float().permute(0, 2, 3, 1), dev), "s2": _f32(sB, dev),
                       "b2": _f32(bB, dev), "wd": None}
                if (p + "downsample.0.weight") in sd:
                    wd = sd[p + "downsample.0.weight"].float()              # [planes, inpl, 1, 1]
                    w3 = torch.zeros(planes, 3, 3, inpl)
                    w3[:, 1, 1, :] = wd[:, :, 0, 0]
                    sD, bD = _bn(sd, p + "downsample.1")
                    blk["wd"], blk["sd"], blk["bd"] = _bf16(w3, dev), _f32(sD, dev), _f32(bD, dev)
                self.blocks.append(blk)
                inpl = planes
        self.decoder = UNetDecoderPlan(sd, dec_prefix, [cfg.in_chans, 64, 64, 128, 256, 512], cfg.n_classes, cfg.patch,
                                       max_batch, dev, decoder_channels=cfg.decoder_channels)
        # workspace (bf16 NHWC)
        P, B, bf = cfg.patch, max_batch, nv.op_dtype()
        self.xn = torch.empty((B, cfg.in_chans, P, P), dtype=torch.float32, device=dev)
        self.f1 = torch.empty((B, P // 2, P // 2, 64), dtype=bf, device=dev)
        self.feat = [torch.empty((B, P // s, P // s, c), dtype=bf, device=dev)
                     for s, c in ((4, 64), (8, 128), (16, 256), (32, 512))]
        big = (P // 4) * (P // 4) * 64
        self.ta = torch.empty(B * big, dtype=bf, device=dev)
        self.tb = torch.empty(B * big, dtype=bf, device=dev)
        self.tc = torch.empty(B * big, dtype=bf, device=dev)
        self.norm = (norm_mean, norm_std)

    def _view(self, buf, n, h, c):
        return buf[:n * h * h * c].view(n, h, h, c)

    def encode_f32(self, x_nchw: torch.Tensor) -> None:
        cfg = self.cfg
        n, P = x_nchw.shape[0], cfg.patch
        assert n <= self.B and x_nchw.shape[1] == cfg.in_chans
        nv.conv7x7s2_bn_relu(x_nchw, self.c1_w_f32, self.c1_s, self.c1_b_f32, self.f1[:n])
        h = P // 4
        cur = self._view(self.ta, n, h, 64)
        nv.maxpool3x3s2(self.f1[:n], cur)
        scratch = [self.ta, self.tb, self.tc]
        cur_buf = 0                                   # index of the scratch buffer holding `cur` (None: a feature map)
        k = 0
        for li, nb in enumerate(cfg.layers):
            for j in range(nb):
                blk = self.blocks[k]
                k += 1
                ho = h // blk["stride"]
                free = [i for i in range(3) if i != cur_buf]
                t = self._view(scratch[free.pop(0)], n, ho, blk["cout"])
                nv.conv3x3(cur, blk["w1"], blk["s1"], blk["b1"], nv.CONV_RELU_BF16, out=t, stride=blk["stride"])
                if blk["wd"] is not None:
                    idn = self._view(scratch[free.pop(0)], n, ho, blk["cout"])
                    nv.conv3x3(cur, blk["wd"], blk["sd"], blk["bd"], nv.CONV_BF16, out=idn, stride=blk["stride"])
                else:
                    idn = cur
                if j == nb - 1:
                    out, out_buf = self.feat[li][:n], None
                elif cur_buf is not None:
                    # conv2 reads t (and idn): cur's buffer is free again; when idn IS cur the epilogue reads and
                    # writes the same element in the same thread, so the in-place residual add is safe
                    out, out_buf = self._view(scratch[cur_buf], n, ho, blk["cout"]), cur_buf
                else:
                    out_buf = free.pop(0)
                    out = self._view(scratch[out_buf], n, ho, blk["cout"])
                nv.conv3x3(t, blk["w2"], blk["s2"], blk["b2"], nv.CONV_ADD_RELU_BF16, out=out, resid=idn)
                cur, cur_buf, h = out, out_buf, ho

    def encode_u8(self, tiles_u8: torch.Tensor, mean: torch.Tensor, std: torch.Tensor) -> None:
        raise nv.NativeError("resnet engine: feed normalised float tiles (fz_gather_tiles_f32) -- see __init__ note")

    def features(self, n: int):
        return [self.f1[:n]] + [f[:n] for f in self.feat]

    def _feats_deep_first(self, n: int):
        return [self.feat[3][:n], self.feat[2][:n], self.feat[1][:n], self.feat[0][:n], self.f1[:n]]

    def decode_logits_nchw(self, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self.decoder.logits_nchw(self._feats_deep_first(n), n, out)

    def decode_argmax_to_raster(self, n, plan, own, raster, margin) -> None:
        self.decoder.argmax_to_raster(self._feats_deep_first(n), n, plan, own, raster, margin)

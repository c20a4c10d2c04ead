"""Oracle (test infrastructure): fp32 eager PyTorch restatement of the model arithmetic.

The reference's model arithmetic lives in two un-vendored third-party packages,
``segmentation-models-pytorch==0.4.0`` (requirements.txt:12) and its transitive ``timm``;
the reference only calls ``smp.create_model`` (flair_hub/models/monotemp_model.py:68-92)
and wires the pieces in ``flair_hub/models/flair_model.py:47-190,357-430,437-547``.  Both
packages are absent from this image, so their published algorithms are restated here
 with the *module tree / state_dict key layout the reference
checkpoints use* (SURVEY.md appendix C):

  encoders.<MOD>.seg_model.model.{stem_0,stem_1,stages_<i>...}      timm FeatureListNet
  encoders.<MOD>.seg_model.{conv1,bn1,layer<k>...}                  smp native ResNetEncoder
  main_decoders.<TASK>.seg_model.decoder.blocks.<k>.conv{1,2}.{0,1}
  main_decoders.<TASK>.seg_model.segmentation_head.0
  fusion_handler.conv_f.<i>

PARITY UNPINNED by the reference (no tests/goldens there); pinned here against HF
``ConvNextV2Model`` and torchvision ``resnet34`` in tests/test_oracle_models.py.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .swin_upernet import SWIN_CFGS, SwinUniversalEncoder, UPerNetDecoder, UPerNetHead, WindowAttention


# --------------------------------------------------------------------------------------
# timm pieces (ConvNeXt-V2)
# --------------------------------------------------------------------------------------
class LayerNorm2d(nn.LayerNorm):
    """timm ``LayerNorm2d``: LayerNorm over the channel dim of an NCHW tensor."""

    def __init__(self, num_channels: int, eps: float = 1e-6):
        super().__init__(num_channels, eps=eps)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


class GlobalResponseNorm(nn.Module):
    """timm ``GlobalResponseNorm`` (channels-last): Gx = ||x||_2 over (H,W);
    Nx = Gx / (mean_c Gx + eps); out = x + bias + weight * (x * Nx)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.zeros(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # (B,H,W,C)
        x_g = x.norm(p=2, dim=(1, 2), keepdim=True)
        x_n = x_g / (x_g.mean(dim=-1, keepdim=True) + self.eps)
        return x + torch.addcmul(self.bias.view(1, 1, 1, -1), self.weight.view(1, 1, 1, -1), x * x_n)


class GlobalResponseNormMlp(nn.Module):
    """timm ``GlobalResponseNormMlp``: fc1 -> GELU(erf) -> GRN -> fc2 on channels-last."""

    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.grn = GlobalResponseNorm(hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.grn(self.act(self.fc1(x))))


class ConvNeXtBlock(nn.Module):
    """timm ``ConvNeXtBlock`` as configured by convnextv2_* (use_grn, no layer-scale,
    conv_mlp=False): dw7x7 -> NHWC -> LN(1e-6) -> MLP(GRN) -> NCHW -> + shortcut."""

    def __init__(self, dim: int):
        super().__init__()
        self.conv_dw = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim, bias=True)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = GlobalResponseNormMlp(dim, 4 * dim)

    def forward(self, x):
        shortcut = x
        x = self.conv_dw(x)
        x = x.permute(0, 2, 3, 1)
        x = self.norm(x)
        x = self.mlp(x)
        x = x.permute(0, 3, 1, 2)
        return x + shortcut


class ConvNeXtStage(nn.Module):
    def __init__(self, in_chs: int, out_chs: int, depth: int, downsample: bool):
        super().__init__()
        if downsample:
            self.downsample = nn.Sequential(LayerNorm2d(in_chs), nn.Conv2d(in_chs, out_chs, kernel_size=2, stride=2))
        else:
            self.downsample = nn.Identity()
        self.blocks = nn.Sequential(*[ConvNeXtBlock(out_chs) for _ in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class ConvNeXtV2Features(nn.Module):
    """timm ``convnextv2_*`` wrapped by ``FeatureListNet(flatten_sequential=True)``:
    children stem_0 (conv4x4 s4), stem_1 (LayerNorm2d), stages_0..3; returns the four
    stage outputs (strides 4/8/16/32).  No final norm in features_only mode."""

    def __init__(self, in_chans: int, depths: Sequence[int], dims: Sequence[int]):
        super().__init__()
        self.stem_0 = nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4)
        self.stem_1 = LayerNorm2d(dims[0])
        prev = dims[0]
        for i, (d, c) in enumerate(zip(depths, dims)):
            setattr(self, f"stages_{i}", ConvNeXtStage(prev, c, d, downsample=i > 0))
            prev = c
        self.num_stages = len(depths)

    def forward(self, x) -> List[torch.Tensor]:
        x = self.stem_1(self.stem_0(x))
        feats = []
        for i in range(self.num_stages):
            x = getattr(self, f"stages_{i}")(x)
            feats.append(x)
        return feats


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


class TimmUniversalEncoder(nn.Module):
    """smp 0.4.0 ``TimmUniversalEncoder`` for a "transformer-style" backbone (first
    feature at stride 4): features = [x, empty(B,0,H/2,W/2), f4, f8, f16, f32] and
    ``out_channels = [C_in, 0, c4, c8, c16, c32]`` -- the 0-channel dummy convention that
    flair_model.py:206-207,302-306,506-518 is written against."""

    def __init__(self, name: str, in_channels: int):
        super().__init__()
        depths, dims = CONVNEXTV2_CFGS[name]
        self.model = ConvNeXtV2Features(in_channels, depths, dims)
        self.out_channels = [in_channels, 0] + list(dims)
        self.output_stride = 32

    def forward(self, x):
        feats = self.model(x)
        b, _, h, w = x.shape
        dummy = torch.empty([b, 0, h // 2, w // 2], dtype=x.dtype, device=x.device)
        return [x, dummy] + feats


# --------------------------------------------------------------------------------------
# smp native ResNet encoder (torchvision ResNet minus fc)
# --------------------------------------------------------------------------------------
class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.relu(out + identity)


class ResNetEncoder(nn.Module):
    """smp ``ResNetEncoder`` (BasicBlock, layers e.g. [3,4,6,3] for resnet34); conv1 is
    rebuilt for ``in_channels != 3`` (smp ``patch_first_conv``; weights come from the
    checkpoint anyway).  features = [x, relu(bn1(conv1 x)), layer1(maxpool .), layer2,
    layer3, layer4]."""

    def __init__(self, in_channels: int, layers: Sequence[int] = (3, 4, 6, 3)):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make_layer(64, layers[0], 1)
        self.layer2 = self._make_layer(128, layers[1], 2)
        self.layer3 = self._make_layer(256, layers[2], 2)
        self.layer4 = self._make_layer(512, layers[3], 2)
        self.out_channels = [in_channels, 64, 64, 128, 256, 512]

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes))
        layers = [BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        layers += [BasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        f0 = x
        f1 = self.relu(self.bn1(self.conv1(x)))
        f2 = self.layer1(self.maxpool(f1))
        f3 = self.layer2(f2)
        f4 = self.layer3(f3)
        f5 = self.layer4(f4)
        return [f0, f1, f2, f3, f4, f5]


# --------------------------------------------------------------------------------------
# smp 0.4.0 U-Net decoder + segmentation head
# --------------------------------------------------------------------------------------
def conv2d_relu(cin: int, cout: int) -> nn.Sequential:
    """smp ``Conv2dReLU(use_batchnorm=True)``: conv3x3 (no bias) + BatchNorm2d + ReLU."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class UnetDecoderBlock(nn.Module):
    def __init__(self, cin: int, cskip: int, cout: int):
        super().__init__()
        self.conv1 = conv2d_relu(cin + cskip, cout)
        self.conv2 = conv2d_relu(cout, cout)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


class UnetDecoder(nn.Module):
    """smp 0.4.0 ``UnetDecoder`` (n_blocks=5, decoder_channels=(256,128,64,32,16),
    batchnorm, no attention, center=Identity)."""

    def __init__(self, encoder_channels: Sequence[int], decoder_channels: Sequence[int] = (256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(encoder_channels)[1:][::-1]
        head = enc[0]
        in_ch = [head] + list(decoder_channels[:-1])
        skip_ch = list(enc[1:]) + [0]
        self.blocks = nn.ModuleList([UnetDecoderBlock(i, s, o) for i, s, o in zip(in_ch, skip_ch, decoder_channels)])

    def forward(self, *features):
        features = features[1:][::-1]
        x = features[0]
        skips = features[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class SegmentationHead(nn.Sequential):
    """smp ``SegmentationHead``: conv k x k (bias) [+ Identity upsampling/activation]."""

    def __init__(self, cin: int, classes: int, kernel_size: int = 3):
        super().__init__(nn.Conv2d(cin, classes, kernel_size, padding=kernel_size // 2))


class DecoderWrapper(nn.Module):
    """monotemp_model.py:7-31."""

    def __init__(self, decoder, segmentation_head):
        super().__init__()
        self.decoder = decoder
        self.segmentation_head = segmentation_head

    def forward(self, *features):
        return self.segmentation_head(self.decoder(*features))


def make_encoder(name: str, in_channels: int) -> nn.Module:
    """What ``smp.create_model(arch, encoder_name=name | 'tu-'+name).encoder`` resolves to
    (monotemp_model.py:67-92)."""
    base = name[3:] if name.startswith("tu-") else name
    if base in CONVNEXTV2_CFGS:
        return TimmUniversalEncoder(base, in_channels)
    if base == "resnet34":
        return ResNetEncoder(in_channels, (3, 4, 6, 3))
    if base == "resnet18":
        return ResNetEncoder(in_channels, (2, 2, 2, 2))
    if base in SWIN_CFGS:
        return SwinUniversalEncoder(base, in_channels)
    raise KeyError(f"oracle: encoder '{name}' not restated")


def make_decoder(arch: str, encoder_channels: Sequence[int], classes: int) -> DecoderWrapper:
    if arch.lower() == "unet":
        return DecoderWrapper(UnetDecoder(encoder_channels), SegmentationHead(16, classes, 3))
    if arch.lower() == "upernet":
        return DecoderWrapper(UPerNetDecoder(encoder_channels), UPerNetHead(64, classes))
    raise KeyError(f"oracle: decoder '{arch}' not restated")


class FLAIRMonotemp(nn.Module):
    """monotemp_model.py:34-97: keeps ``.seg_model`` = encoder, or DecoderWrapper."""

    def __init__(self, arch: str, channels: int, classes: int, return_type: str):
        super().__init__()
        encoder, decoder = arch.split("-")[0], arch.split("-")[1]
        enc = make_encoder(encoder, channels)
        if return_type == "encoder":
            self.seg_model = enc
        else:
            self.seg_model = make_decoder(decoder, enc.out_channels, classes)


class FusionHandler(nn.Module):
    """flair_model.py:437-547 (mono-only cases: 1 key passthrough, >=2 keys concat + 1x1)."""

    def __init__(self, backbones_channels, target_fused_channels):
        super().__init__()
        t = list(target_fused_channels)
        if len(t) > 2 and (t[0] == 0 or t[1] == 0):
            t = t[2:]
        self.conv_f = nn.ModuleList([nn.Conv2d(i, o, kernel_size=1) for i, o in zip(backbones_channels, t)])

    def forward(self, feature_maps: Dict[str, List[torch.Tensor]], target: List[torch.Tensor]):
        keys = list(feature_maps.keys())
        if len(keys) == 1:
            return feature_maps[keys[0]]
        shapes = [fm.shape for fm in target]
        dummy = None
        if shapes[0][1] == 0 or shapes[1][1] == 0:
            shapes = shapes[2:]
            dummy = target[:2]
        aligned = []
        for mod in keys:
            fms = feature_maps[mod]
            if fms[0].shape[1] == 0 or fms[1].shape[1] == 0:
                fms = fms[2:]
            if len(fms) != len(shapes):
                fms = [fms[0]] * (len(shapes) - len(fms)) + fms
            res = []
            for fm, t in zip(fms, shapes):
                if fm.shape[-1] != t[-1] or fm.shape[-2] != t[-2]:
                    fm = F.interpolate(fm, size=(t[-2], t[-1]), mode="bilinear", align_corners=False)
                res.append(fm)
            aligned.append(res)
        stacked = [torch.cat(f, dim=1) for f in zip(*aligned)]
        out = [c(f) for c, f in zip(self.conv_f, stacked)]
        if dummy is not None:
            out = list(dummy) + out
        return out


class FlairHubOracle(nn.Module):
    """flair_model.py:16-430 restricted to mono-temporal modalities (the only ones in
    BASELINE.json's configs).  ``forward(batch) -> (logits_tasks, logits_aux)``."""

    MONO_KEYS = ["AERIAL_RGBI", "AERIAL-RLT_PAN", "DEM_ELEV", "SPOT_RGBI"]

    def __init__(self, arch: str, modalities: Dict[str, int], tasks: Dict[str, int]):
        """modalities: {MOD: n_input_channels} (active ones); tasks: {TASK: n_classes}."""
        super().__init__()
        self.arch = arch
        self.labels = list(tasks.keys())
        self.task_nclasses = sum(tasks.values())
        self.encoders = nn.ModuleDict()
        for mod in self.MONO_KEYS:
            if mod in modalities:
                self.encoders[mod] = FLAIRMonotemp(arch, modalities[mod], self.task_nclasses, "encoder")
        chans = []
        for mod in self.encoders:
            oc = self.encoders[mod].seg_model.out_channels
            chans.append(list(oc[2:]) if len(oc) > 2 and (oc[0] == 0 or oc[1] == 0) else list(oc))
        total = [sum(x) for x in zip(*chans)]
        target = next(iter(self.encoders.values())).seg_model.out_channels
        self.fusion_handler = FusionHandler(total, target)
        self.main_decoders = nn.ModuleDict(
            {t: FLAIRMonotemp(arch, 1, n, "decoder") for t, n in tasks.items()})
        self.aux_decoders = nn.ModuleDict()

    @staticmethod
    def modality_dropout(feature_maps: Dict[str, list], probs: Dict[str, float]) -> Dict[str, list]:
        """flair_model.py:330-354: per modality one ``torch.rand(1)`` against its probability; a dropped modality's maps are
        replaced by xavier-uniform noise of the same shapes (fresh tensors: nothing upstream receives a gradient)."""
        import warnings
        for key in feature_maps.keys():
            if torch.rand(1).item() < probs[key]:
                noise = []
                for t in feature_maps[key]:
                    n = torch.empty_like(t)
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")              # zero-channel dummy map: initialisation is a no-op
                        nn.init.xavier_uniform_(n)
                    noise.append(n)
                feature_maps[key] = noise
        return feature_maps

    def forward(self, batch: Dict[str, torch.Tensor], apply_mod_dropout: bool = False):
        img_size = batch[self.labels[0]].shape[-1]
        fmaps = {mod: enc.seg_model(batch[mod]) for mod, enc in self.encoders.items()}
        if apply_mod_dropout and len(self.encoders) > 1:             # flair_model.py:406-408: the probabilities are drawn too
            import random
            fmaps = self.modality_dropout(fmaps, {key: random.uniform(0, 1) for key in fmaps.keys()})
        first = next(iter(self.encoders.keys()))
        fused = self.fusion_handler(fmaps, fmaps[first])
        logits = {}
        for t in self.labels:
            y = self.main_decoders[t].seg_model(*fused)
            logits[t] = F.interpolate(y, size=img_size, mode="bilinear", align_corners=False)
        return logits, {}


@torch.no_grad()
def randomize_(module: nn.Module, seed: int = 2025, bf16_exact: bool = True) -> None:
    """Randomise EVERY parameter and buffer (SURVEY.md appendix C caveat: default inits
    leave GRN, biases, BN statistics inert) with magnitudes that keep activations O(1)
    through the network.  ``bf16_exact`` snaps values to bf16-representable numbers so a
    bf16 kernel sees the same weights as the fp32 oracle."""
    g = torch.Generator().manual_seed(seed)

    def rn(shape, std):
        return torch.randn(shape, generator=g) * std

    def ru(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    for name, mod in module.named_modules():
        if isinstance(mod, (nn.Conv2d, nn.Linear)):
            w = mod.weight
            fan_in = w[0].numel()
            w.copy_(rn(w.shape, (2.0 / fan_in) ** 0.5 if not isinstance(mod, nn.Linear) else (1.0 / fan_in) ** 0.5))
            if name.endswith("mlp.fc2"):
                w.mul_(0.5)
            if mod.bias is not None:
                mod.bias.copy_(rn(mod.bias.shape, 0.1))
        elif isinstance(mod, nn.BatchNorm2d):
            mod.weight.copy_(ru(mod.weight.shape, 0.5, 1.5))
            mod.bias.copy_(rn(mod.bias.shape, 0.1))
            mod.running_mean.copy_(rn(mod.running_mean.shape, 0.1))
            mod.running_var.copy_(ru(mod.running_var.shape, 0.5, 1.5))
        elif isinstance(mod, nn.LayerNorm):
            mod.weight.copy_(ru(mod.weight.shape, 0.5, 1.5))
            mod.bias.copy_(rn(mod.bias.shape, 0.1))
        elif isinstance(mod, GlobalResponseNorm):
            mod.weight.copy_(rn(mod.weight.shape, 0.5))
            mod.bias.copy_(rn(mod.bias.shape, 0.1))
        elif isinstance(mod, WindowAttention):
            mod.relative_position_bias_table.copy_(rn(mod.relative_position_bias_table.shape, 0.5))
    if bf16_exact:
        for p in list(module.parameters()) + [b for b in module.buffers() if b.dtype.is_floating_point]:
            p.copy_(p.to(torch.bfloat16).to(p.dtype))

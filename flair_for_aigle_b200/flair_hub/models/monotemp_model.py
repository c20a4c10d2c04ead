"""Mono-temporal encoder / decoder descriptions: drop-in for flair_hub/models/monotemp_model.py.

The reference's ``FLAIR_Monotemp`` (monotemp_model.py:34-97) asks ``smp.create_model`` for a
full segmentation model and keeps either ``.encoder`` or ``DecoderWrapper(decoder, head)``.
Here the "model" is a *parameter specification*: the exact state_dict keys and shapes that smp
0.4.0 / timm would create, which ``FLAIR_HUB_Model`` registers as plain
tensors and the sm_100a engine packs for its kernels.  No torch.nn compute modules are built.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

from ...engine.convnext_unet import CONVNEXTV2_CFGS
from ...engine.resnet_unet import RESNET_LAYERS
from ...engine.swin_upernet import SWIN_CFGS

# kind drives the random initialisation only
Spec = "OrderedDict[str, Tuple[Tuple[int, ...], str]]"

UNET_DECODER_CHANNELS = (256, 128, 64, 32, 16)


def split_arch(arch: str) -> Tuple[str, str]:
    """monotemp_model.py:64-65: "<encoder>-<decoder>"."""
    parts = arch.split("-")
    return parts[0], parts[1]


def resolve_encoder(name: str) -> str:
    """The smp lookup of monotemp_model.py:67-92: native encoder name first, then 'tu-'+name."""
    base = name[3:] if name.startswith("tu-") else name
    if base in CONVNEXTV2_CFGS or base in RESNET_LAYERS or base in SWIN_CFGS:
        return base
    raise KeyError(
        f"encoder '{name}' has no sm_100a execution plan yet (available: "
        f"{sorted(CONVNEXTV2_CFGS) + sorted(RESNET_LAYERS) + sorted(SWIN_CFGS)}); there is no PyTorch fallback")


def encoder_out_channels(name: str, in_channels: int) -> List[int]:
    """smp ``encoder.out_channels``; timm-universal "transformer style": [C_in, 0, c4, c8, c16, c32]."""
    base = resolve_encoder(name)
    if base in RESNET_LAYERS:
        return [in_channels, 64, 64, 128, 256, 512]          # smp native ResNetEncoder
    if base in SWIN_CFGS:
        dim = SWIN_CFGS[base][0]
        return [in_channels, 0] + [dim * 2 ** i for i in range(4)]
    _, dims = CONVNEXTV2_CFGS[base]
    return [in_channels, 0] + list(dims)


def convnextv2_encoder_spec(name: str, in_channels: int):
    """Keys below ``encoders.<MOD>.seg_model.`` (smp TimmUniversalEncoder.model = timm
    FeatureListNet with flatten_sequential=True)."""
    depths, dims = CONVNEXTV2_CFGS[resolve_encoder(name)]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["model.stem_0.weight"] = ((dims[0], in_channels, 4, 4), "conv")
    s["model.stem_0.bias"] = ((dims[0],), "bias")
    s["model.stem_1.weight"] = ((dims[0],), "norm_w")
    s["model.stem_1.bias"] = ((dims[0],), "bias")
    prev = dims[0]
    for i, (d, c) in enumerate(zip(depths, dims)):
        p = f"model.stages_{i}."
        if i > 0:
            s[p + "downsample.0.weight"] = ((prev,), "norm_w")
            s[p + "downsample.0.bias"] = ((prev,), "bias")
            s[p + "downsample.1.weight"] = ((c, prev, 2, 2), "conv")
            s[p + "downsample.1.bias"] = ((c,), "bias")
        for j in range(d):
            b = p + f"blocks.{j}."
            s[b + "conv_dw.weight"] = ((c, 1, 7, 7), "conv")
            s[b + "conv_dw.bias"] = ((c,), "bias")
            s[b + "norm.weight"] = ((c,), "norm_w")
            s[b + "norm.bias"] = ((c,), "bias")
            s[b + "mlp.fc1.weight"] = ((4 * c, c), "linear")
            s[b + "mlp.fc1.bias"] = ((4 * c,), "bias")
            s[b + "mlp.grn.weight"] = ((4 * c,), "grn")
            s[b + "mlp.grn.bias"] = ((4 * c,), "grn")
            s[b + "mlp.fc2.weight"] = ((c, 4 * c), "linear")
            s[b + "mlp.fc2.bias"] = ((c,), "bias")
        prev = c
    return s


def unet_decoder_spec(encoder_channels: Sequence[int], classes: int,
                      decoder_channels: Sequence[int] = UNET_DECODER_CHANNELS):
    """Keys below ``main_decoders.<TASK>.seg_model.`` (smp 0.4.0 UnetDecoder + SegmentationHead)."""
    enc = list(encoder_channels)[1:][::-1]
    in_ch = [enc[0]] + list(decoder_channels[:-1])
    skip_ch = list(enc[1:]) + [0]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    for k, (ci, cs, co) in enumerate(zip(in_ch, skip_ch, decoder_channels)):
        for name, cin in (("conv1", ci + cs), ("conv2", co)):
            p = f"decoder.blocks.{k}.{name}."
            s[p + "0.weight"] = ((co, cin, 3, 3), "conv_relu")
            s[p + "1.weight"] = ((co,), "norm_w")
            s[p + "1.bias"] = ((co,), "bias")
            s[p + "1.running_mean"] = ((co,), "bn_mean")
            s[p + "1.running_var"] = ((co,), "bn_var")
            s[p + "1.num_batches_tracked"] = ((), "bn_count")
    s["segmentation_head.0.weight"] = ((classes, decoder_channels[-1], 3, 3), "head")
    s["segmentation_head.0.bias"] = ((classes,), "bias")
    return s


def _bn_spec(s, prefix: str, c: int):
    s[prefix + ".weight"] = ((c,), "norm_w")
    s[prefix + ".bias"] = ((c,), "bias")
    s[prefix + ".running_mean"] = ((c,), "bn_mean")
    s[prefix + ".running_var"] = ((c,), "bn_var")
    s[prefix + ".num_batches_tracked"] = ((), "bn_count")


def resnet_encoder_spec(name: str, in_channels: int):
    """Keys below ``encoders.<MOD>.seg_model.`` for smp's native ResNetEncoder (torchvision ResNet, BasicBlock,
    no fc).  (If smp resolved the name through timm instead the keys would carry a ``model.`` prefix; SURVEY.md
    appendix A -- the checkpoint loader is tolerant either way.)"""
    layers = RESNET_LAYERS[resolve_encoder(name)]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["conv1.weight"] = ((64, in_channels, 7, 7), "conv_relu")
    _bn_spec(s, "bn1", 64)
    inpl = 64
    for li, (nb, planes) in enumerate(zip(layers, (64, 128, 256, 512))):
        for j in range(nb):
            stride = 2 if (j == 0 and li > 0) else 1
            p = f"layer{li + 1}.{j}."
            s[p + "conv1.weight"] = ((planes, inpl, 3, 3), "conv_relu")
            _bn_spec(s, p + "bn1", planes)
            s[p + "conv2.weight"] = ((planes, planes, 3, 3), "conv_relu")
            _bn_spec(s, p + "bn2", planes)
            if stride != 1 or inpl != planes:
                s[p + "downsample.0.weight"] = ((planes, inpl, 1, 1), "conv_relu")
                _bn_spec(s, p + "downsample.1", planes)
            inpl = planes
    return s


def swin_encoder_spec(name: str, in_channels: int, img_size: int = 512):
    """Keys below ``encoders.<MOD>.seg_model.`` for a timm Swin behind smp's TimmUniversalEncoder: ``model`` is
    timm's FeatureGetterNet, which keeps the backbone as ``.model`` (head and final norm pruned).  The
    ``relative_position_index`` / ``attn_mask`` buffers are non-persistent in timm and carry no keys."""
    dim, depths, heads, window = SWIN_CFGS[resolve_encoder(name)]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    p = "model.model."
    s[p + "patch_embed.proj.weight"] = ((dim, in_channels, 4, 4), "conv")
    s[p + "patch_embed.proj.bias"] = ((dim,), "bias")
    s[p + "patch_embed.norm.weight"] = ((dim,), "norm_w")
    s[p + "patch_embed.norm.bias"] = ((dim,), "bias")
    res = img_size // 4
    prev = dim
    for i, (d, nh) in enumerate(zip(depths, heads)):
        c = dim * 2 ** i
        L = p + f"layers.{i}."
        if i > 0:
            res //= 2
            s[L + "downsample.norm.weight"] = ((4 * prev,), "norm_w")
            s[L + "downsample.norm.bias"] = ((4 * prev,), "bias")
            s[L + "downsample.reduction.weight"] = ((c, 4 * prev), "linear")
        w_eff = res if res <= window else window
        for j in range(d):
            b = L + f"blocks.{j}."
            s[b + "norm1.weight"] = ((c,), "norm_w")
            s[b + "norm1.bias"] = ((c,), "bias")
            s[b + "attn.relative_position_bias_table"] = (((2 * w_eff - 1) ** 2, nh), "table")
            s[b + "attn.qkv.weight"] = ((3 * c, c), "linear")
            s[b + "attn.qkv.bias"] = ((3 * c,), "bias")
            s[b + "attn.proj.weight"] = ((c, c), "linear")
            s[b + "attn.proj.bias"] = ((c,), "bias")
            s[b + "norm2.weight"] = ((c,), "norm_w")
            s[b + "norm2.bias"] = ((c,), "bias")
            s[b + "mlp.fc1.weight"] = ((4 * c, c), "linear")
            s[b + "mlp.fc1.bias"] = ((4 * c,), "bias")
            s[b + "mlp.fc2.weight"] = ((c, 4 * c), "linear")
            s[b + "mlp.fc2.bias"] = ((c,), "bias")
        prev = c
    return s


def _conv_bn_spec(s, prefix: str, cin: int, cout: int, k: int):
    s[prefix + ".0.weight"] = ((cout, cin, k, k), "conv_relu")
    _bn_spec(s, prefix + ".1", cout)


def upernet_decoder_spec(encoder_channels: Sequence[int], classes: int, pyramid: int = 256, seg: int = 64):
    """Keys below ``main_decoders.<TASK>.seg_model.`` (smp 0.4.0 UPerNetDecoder + SegmentationHead(k=1, up=4)).
    smp builds one FPN block per encoder channel after the deepest, including one for the input-resolution
    feature that forward() never reaches; the 0-channel entry is an Identity (no keys)."""
    enc = list(encoder_channels)[::-1]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    for k in range(4):
        _conv_bn_spec(s, f"decoder.psp.blocks.{k}.1", enc[0], enc[0] // 4, 1)
    _conv_bn_spec(s, "decoder.psp.out_conv", 2 * enc[0], pyramid, 1)
    for k, ch in enumerate(enc[1:]):
        if ch != 0:
            _conv_bn_spec(s, f"decoder.fpn_stages.{k}.skip_conv", ch, pyramid, 1)
    _conv_bn_spec(s, "decoder.fpn_bottleneck", (len(enc) - 1) * pyramid, seg, 3)
    s["segmentation_head.0.weight"] = ((classes, seg, 1, 1), "head")
    s["segmentation_head.0.bias"] = ((classes,), "bias")
    return s


def encoder_family(name: str) -> str:
    base = resolve_encoder(name)
    return "resnet" if base in RESNET_LAYERS else ("swin" if base in SWIN_CFGS else "convnextv2")


def encoder_spec(arch: str, in_channels: int):
    enc, _ = split_arch(arch)
    if encoder_family(enc) == "resnet":
        return resnet_encoder_spec(enc, in_channels)
    if encoder_family(enc) == "swin":
        return swin_encoder_spec(enc, in_channels)
    return convnextv2_encoder_spec(enc, in_channels)


def decoder_spec(arch: str, in_channels: int, classes: int):
    enc, dec = split_arch(arch)
    if dec.lower() == "upernet":
        return upernet_decoder_spec(encoder_out_channels(enc, in_channels), classes)
    if dec.lower() != "unet":
        raise KeyError(f"decoder '{dec}' has no sm_100a execution plan yet (available: unet, upernet); "
                       "no PyTorch fallback")
    return unet_decoder_spec(encoder_out_channels(enc, in_channels), classes)

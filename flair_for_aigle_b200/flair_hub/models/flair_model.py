"""``FLAIR_HUB_Model``: drop-in for flair_hub/models/flair_model.py:16-430 on the zonal hot path.

Same constructor (``config``, ``img_input_sizes``), same attributes the callers read
(``task_nclasses``, ``mono_keys`` ...), same ``state_dict`` key layout (so the reference's
checkpoints load through ``load_checkpoint``), same call convention::

    logits_tasks, logits_aux = model(inputs)      # inputs[<MOD>]: (B,C,H,W) fp32 normalised

but no torch.nn compute: parameters are plain registered tensors, and ``forward`` runs the
hand-written sm_100a kernels through the C ABI (engine/convnext_unet.py).  On a machine without
a CUDA device, or with an architecture that has no execution plan, it raises -- there is no
PyTorch/CPU fallback.

In scope: one mono-temporal modality (FusionHandler case 1,
flair_model.py:488-490), U-Net decoder.  Out of scope and rejected loudly: Sentinel time-series
encoders (UTAE), auxiliary decoders (dead code in the reference: their loss is identically zero); modality dropout lives in
the training engine (``draw_modality_dropout`` below).
"""
from __future__ import annotations

import logging
import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from ... import native as nv
from ...engine.convnext_unet import CONVNEXTV2_CFGS, ConvNeXtCfg, ConvNeXtV2UNetEngine
from ...engine.fusion import FusedEncodersUNet
from ...engine.resnet_unet import RESNET_LAYERS, ResNetCfg, ResNetUNetEngine
from ...engine.swin_upernet import SWIN_CFGS, SwinCfg, SwinUPerNetEngine
from . import monotemp_model as mm

logger = logging.getLogger(__name__)


def draw_modality_dropout(feature_shapes: Dict[str, List[tuple]], device, dtype=torch.float32) -> Dict[str, List[torch.Tensor]]:
    """The random part of the reference's training-time modality dropout (flair_model.py:406-408 and :330-354), draw for draw:
    one ``random.uniform(0, 1)`` per modality in dict order (the dropout probability of THIS call -- the configured values
    only switch the feature on, tasks_module.py:59-61), then per modality one ``torch.rand(1)`` against it, and for a dropped
    modality every feature map of its list replaced by ``nn.init.xavier_uniform_`` noise of the same (B,C,h,w) shape (a
    zero-channel dummy map draws nothing).  -> {dropped modality: [noise tensors, NCHW]}.  torch's own generators are used on
    purpose: with the same seeds the draws equal the reference's (tests/test_reference_pin.py)."""
    import random
    import warnings
    probs = {key: random.uniform(0, 1) for key in feature_shapes}
    dropped: Dict[str, List[torch.Tensor]] = {}
    for key, shapes in feature_shapes.items():
        if torch.rand(1).item() < probs[key]:
            maps = []
            for shape in shapes:
                t = torch.empty(tuple(shape), device=device, dtype=dtype)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")          # "Initializing zero-element tensors is a no-op"
                    nn.init.xavier_uniform_(t)
                maps.append(t)
            dropped[key] = maps
    return dropped


class _Node(nn.Module):
    """Parameter container (no forward)."""


def _register(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool) -> None:
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, _Node())
        node = node._modules[p]
    if buffer:
        node.register_buffer(parts[-1], tensor)
    else:
        node.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _init_tensor(shape, kind: str, gen: torch.Generator) -> torch.Tensor:
    if kind in ("conv", "linear", "table"):
        return torch.nn.init.trunc_normal_(torch.empty(shape), std=0.02, generator=gen)   # timm default
    if kind == "conv_relu":
        fan_in = shape[1] * shape[2] * shape[3]
        bound = math.sqrt(6.0 / fan_in)                                                   # smp: kaiming_uniform
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound
    if kind == "head":
        fan_in, fan_out = shape[1] * shape[2] * shape[3], shape[0] * shape[2] * shape[3]
        bound = math.sqrt(6.0 / (fan_in + fan_out))                                       # smp: xavier_uniform
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound
    if kind in ("norm_w", "bn_var"):
        return torch.ones(shape)
    if kind == "bn_count":
        return torch.zeros(shape, dtype=torch.long)
    return torch.zeros(shape)  # bias, grn, bn_mean


class FLAIR_HUB_Model(nn.Module):
    mono_keys = ['AERIAL_RGBI', 'AERIAL-RLT_PAN', 'DEM_ELEV', 'SPOT_RGBI']
    multi_keys = ['SENTINEL2_TS', 'SENTINEL1-ASC_TS', 'SENTINEL1-DESC_TS']

    def __init__(self, config: dict, img_input_sizes: dict, max_batch: int = 16):
        super().__init__()
        self.config = config
        self.img_input_sizes = img_input_sizes
        self.max_batch = max_batch
        inputs = config['modalities']['inputs']

        self.aux_losses = {mod: loss for mod, loss in config['modalities'].get('aux_loss', {}).items()
                           if loss and inputs.get(mod, False)}
        if self.aux_losses:
            raise NotImplementedError("auxiliary decoders are outside the zonal hot path (SURVEY.md section 8)")
        if any(inputs.get(k, False) for k in self.multi_keys):
            raise NotImplementedError("Sentinel time-series encoders (UTAE) are outside the zonal hot path")

        self.tasks = len(config['labels'])
        self.task_nclasses = sum(len(config['labels_configs'][label]['value_name']) for label in config['labels'])
        # flair_model.py:69-87
        self.channels_dict = {
            mod: (1 if mod in ['AERIAL-RLT_PAN', 'DEM_ELEV']
                  else (len(config['modalities']['inputs_channels'][mod])
                        if mod in config['modalities']['inputs_channels'] else 0))
            for mod in inputs
        }
        if inputs.get('DEM_ELEV', False):
            pp = config['modalities']['pre_processings']
            self.channels_dict['DEM_ELEV'] = 1 if pp['calc_elevation'] and not pp['calc_elevation_stack_dsm'] else 2

        self.arch = config['models']['monotemp_model']['arch']
        self.active_mono = [m for m in self.mono_keys if inputs.get(m, False)]
        if len(self.active_mono) < 1:
            raise NotImplementedError("no mono-temporal modality is active")
        enc_name, _ = mm.split_arch(self.arch)
        self.encoder_name = mm.resolve_encoder(enc_name)
        if len(self.active_mono) > 1 and self.encoder_name not in CONVNEXTV2_CFGS:
            raise NotImplementedError(
                f"{len(self.active_mono)} mono-temporal modalities: FusionHandler's concat + 1x1 path has an sm_100a "
                "plan for the convnextv2_*-unet family only")

        gen = torch.Generator().manual_seed(int(config.get('seed', 2025)))
        self.encoders = _Node()
        self.fusion_handler = _Node()
        self.main_decoders = _Node()
        self.aux_decoders = _Node()
        for mod in self.active_mono:
            for k, (shape, kind) in mm.encoder_spec(self.arch, self.channels_dict[mod]).items():
                _register(self.encoders, f"{mod}.seg_model.{k}", _init_tensor(shape, kind, gen),
                          buffer=kind.startswith("bn_"))
        first = self.active_mono[0]
        # flair_model.py:141-149 / :466-471: FusionHandler.conv_f is always built (one 1x1 conv per
        # stage), also when a single modality makes it a pass-through; checkpoints carry its weights
        oc = mm.encoder_out_channels(self.encoder_name, self.channels_dict[first])
        stage_ch = oc[2:] if len(oc) > 2 and (oc[0] == 0 or oc[1] == 0) else oc
        for i, c in enumerate(stage_ch):
            tot = c * len(self.active_mono)
            _register(self.fusion_handler, f"conv_f.{i}.weight", _init_tensor((c, tot, 1, 1), "conv", gen), False)
            _register(self.fusion_handler, f"conv_f.{i}.bias", _init_tensor((c,), "bias", gen), False)
        # flair_model.py:151-166: decoders are built with channels=1; only the encoder's
        # out_channels[2:] (independent of the input channel count) reach the decoder
        for task in config['labels']:
            ncls = len(config['labels_configs'][task]['value_name'])
            for k, (shape, kind) in mm.decoder_spec(self.arch, 1, ncls).items():
                _register(self.main_decoders, f"{task}.seg_model.{k}", _init_tensor(shape, kind, gen),
                          buffer=kind.startswith("bn_"))
        self._engines: Dict[str, ConvNeXtV2UNetEngine] = {}
        self._norm = self._normalisation(first)

    # -------------------------------------------------------------------------------- helpers
    def _normalisation(self, mod: str):
        """(means, stds) the fused uint8 stem folds in: every normalisation type of norm.py, for 8-bit imagery (the only
        raster type the fused feeder reads)."""
        from ...flair_zonal_detection.dataset import normalization_affine
        return normalization_affine(self.config['modalities'].get(mod, {}).get('normalization'), self.channels_dict[mod],
                                    np.uint8)

    def _device(self) -> torch.device:
        return next(self.parameters()).device

    def invalidate(self) -> None:
        """Drop packed weights (called after load_state_dict / .to())."""
        self._engines = {}

    def _apply(self, fn, *a, **k):
        self._engines = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engines = {}
        return super().load_state_dict(*a, **k)

    def engine(self, task: Optional[str] = None, max_batch: Optional[int] = None):
        """Packed-weight execution plan for ``task`` on the parameters' device."""
        task = task or self.config['labels'][0]
        mb = max_batch or self.max_batch
        key = f"{task}:{mb}"
        if key not in self._engines and len(self.active_mono) > 1:
            # flair_model.py:376 per modality, :410 FusionHandler (concat + conv_f), :417-419 decoder
            dev = self._device()
            if dev.type != "cuda":
                raise nv.NativeError("FLAIR_HUB_Model runs on hand-written sm_100a kernels only (no CPU fallback)")
            ncls = len(self.config['labels_configs'][task]['value_name'])
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            depths, dims = CONVNEXTV2_CFGS[self.encoder_name]
            encs = {}
            for mod in self.active_mono:
                cfg = ConvNeXtCfg(depths=depths, dims=dims, in_chans=self.channels_dict[mod], n_classes=ncls,
                                  patch=int(self.img_input_sizes[mod]))
                encs[mod] = ConvNeXtV2UNetEngine(sd, f"encoders.{mod}.seg_model.model.",
                                                 f"main_decoders.{task}.seg_model.", cfg, dev, max_batch=mb)
            self._engines = {k: e for k, e in self._engines.items() if k.endswith(f":{mb}")}
            self._engines[key] = FusedEncodersUNet(encs, sd, "fusion_handler.")
        if key not in self._engines:
            dev = self._device()
            if dev.type != "cuda":
                raise nv.NativeError(
                    "FLAIR_HUB_Model runs on hand-written sm_100a kernels only: move it to a CUDA device "
                    "(`.to('cuda')`); there is no CPU fallback")
            mod = self.active_mono[0]
            ncls = len(self.config['labels_configs'][task]['value_name'])
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            mean, std = self._norm if self._norm else (None, None)
            self._engines = {k: e for k, e in self._engines.items() if k.endswith(f":{mb}")}
            if self.encoder_name in RESNET_LAYERS:
                cfg = ResNetCfg(layers=RESNET_LAYERS[self.encoder_name], in_chans=self.channels_dict[mod],
                                n_classes=ncls, patch=int(self.img_input_sizes[mod]))
                self._engines[key] = ResNetUNetEngine(sd, f"encoders.{mod}.seg_model.",
                                                      f"main_decoders.{task}.seg_model.", cfg, dev, max_batch=mb,
                                                      norm_mean=mean, norm_std=std)
            elif self.encoder_name in SWIN_CFGS:
                dim, depths, heads, window = SWIN_CFGS[self.encoder_name]
                cfg = SwinCfg(embed_dim=dim, depths=depths, heads=heads, window=window,
                              in_chans=self.channels_dict[mod], n_classes=ncls, patch=int(self.img_input_sizes[mod]))
                self._engines[key] = SwinUPerNetEngine(sd, f"encoders.{mod}.seg_model.model.model.",
                                                       f"main_decoders.{task}.seg_model.", cfg, dev, max_batch=mb,
                                                       norm_mean=mean, norm_std=std)
            else:
                depths, dims = CONVNEXTV2_CFGS[self.encoder_name]
                cfg = ConvNeXtCfg(depths=depths, dims=dims, in_chans=self.channels_dict[mod], n_classes=ncls,
                                  patch=int(self.img_input_sizes[mod]))
                self._engines[key] = ConvNeXtV2UNetEngine(
                    sd, f"encoders.{mod}.seg_model.model.", f"main_decoders.{task}.seg_model.", cfg, dev,
                    max_batch=mb, norm_mean=mean, norm_std=std)
        return self._engines[key]

    # -------------------------------------------------------------------------------- forward
    @torch.no_grad()
    def _forward_fused(self, batch: dict):
        """flair_model.py:357-430 with >= 2 mono modalities (FusionHandler concat + 1x1, :503-547)."""
        labels = self.config['labels']
        x0 = batch[self.active_mono[0]]
        img_size = batch[labels[0]].shape[-1] if labels[0] in batch else x0.shape[-1]
        if img_size != x0.shape[-1]:
            raise NotImplementedError("final bilinear resize (flair_model.py:327) other than the identity is not built")
        logits_tasks = {}
        for task in labels:
            eng = self.engine(task)
            n = x0.shape[0]
            out = torch.empty((n, eng.cfg.n_classes, x0.shape[-2], x0.shape[-1]), dtype=torch.float32, device=x0.device)
            for s in range(0, n, eng.B):
                e = min(s + eng.B, n)
                eng.encode({m: batch[m][s:e] for m in self.active_mono})
                eng.decode_logits_nchw(e - s, out=out[s:e])
            logits_tasks[task] = out
        return logits_tasks, {}

    @torch.no_grad()
    def forward(self, batch: dict, apply_mod_dropout: bool = False):
        """flair_model.py:357-430 for one mono modality: returns ({task: (B,n_cls,H,W) fp32}, {})."""
        if apply_mod_dropout:
            raise NotImplementedError("modality dropout belongs to the training step: SegmentationTask.training_step / "
                                      "engine.train_step.ConvNeXtUNetTrainer(mod_dropout=True) apply it (draw_modality_dropout)")
        if len(self.active_mono) > 1:
            return self._forward_fused(batch)
        mod = self.active_mono[0]
        x = batch[mod]
        if not x.is_cuda:
            raise nv.NativeError("inputs must be CUDA tensors (no CPU fallback)")
        x = x.contiguous().float()
        labels = self.config['labels']
        img_size = batch[labels[0]].shape[-1] if labels[0] in batch else x.shape[-1]
        if img_size != x.shape[-1]:
            raise NotImplementedError(
                f"label size {img_size} != input size {x.shape[-1]}: the final bilinear resize "
                "(flair_model.py:327) is the identity in every zonal configuration and is not implemented otherwise")
        logits_tasks = {}
        for task in labels:
            eng = self.engine(task)
            n = x.shape[0]
            out = torch.empty((n, eng.cfg.n_classes, x.shape[-2], x.shape[-1]), dtype=torch.float32, device=x.device)
            for s in range(0, n, eng.B):
                e = min(s + eng.B, n)
                eng.encode_f32(x[s:e])
                eng.decode_logits_nchw(e - s, out=out[s:e])
            logits_tasks[task] = out
        return logits_tasks, {}

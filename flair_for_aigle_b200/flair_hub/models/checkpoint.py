"""Checkpoint loading: drop-in for ``load_checkpoint`` of flair_hub/models/checkpoint.py:176-290.

The reference patches the checkpoint's tensors until ``load_state_dict(strict=False)`` cannot fail, in a fixed
order that matters because every re-initialisation draws from torch's global RNG.  Here the same behaviour is
a short list of REPAIR PASSES over the loaded tensors, each pass a pure function of (checkpoint tensors, module
tensors) that returns the replacements it wants; ``load_checkpoint`` applies them in the reference's order, so
that under the same ``torch.manual_seed`` the loaded module is bit-identical to the reference's
(``tests/test_reference_pin.py::test_load_checkpoint_*`` runs both on crafted checkpoints).

  pass                    reference lines   rule
  ----------------------  ----------------  ---------------------------------------------------------------
  read                    :206-212          ``.safetensors`` -> safetensors, else ``torch.load(path)["state_dict"]``
                                            (a bare state dict is accepted too); bad path -> ``SystemExit`` unless
                                            ``exit_on_fail=False`` (:200-204)
  lightning prefix        :134-173          drop a leading ``model.`` from checkpoint keys iff some checkpoint key
                                            has it and NO module key has it
  task heads              :224-241,:87-131  per task, the head ``main_decoders.<task>.seg_model.segmentation_head.0``
                                            is looked up with and without ``model.``; a missing head or one with
                                            another class count is re-initialised (weight, then bias)
  auxiliary heads         :243-250          same for module keys under ``model.aux_decoders.``
  criterion weights       :252-259          a class-weight vector of another length takes the module's
  shape mismatches        :261-274          ``relative_position_bias_table`` -> bicubic resize of the square table
                                            (:33-56); anything else -> re-initialised
  load                    :278              ``load_state_dict(strict=False)``

Re-initialisation (:11-31) goes by NAME: ``weight`` in the key -> Xavier-uniform (a 1-D tensor therefore raises
``ValueError``, as in the reference); ``bias`` in the key -> zeros (this catches ``relative_position_bias_table``
too); any other tensor (BatchNorm statistics) is left as ``torch.empty_like`` by the reference, i.e. undefined --
zeros here.  Deviation, deliberate: the reference's ``@rank_zero_only`` makes every rank but 0 skip the load and
rely on DDP's broadcast; the zonal strips have no DDP wrapper, so every rank loads.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Iterator, List, Optional, Set, Tuple

import torch
import torch.nn as nn

logger = logging.getLogger(__name__)

Tensors = Dict[str, torch.Tensor]
HEAD = "seg_model.segmentation_head.0.weight"


@dataclass
class LoadReport:
    matched_tasks: Set[str] = field(default_factory=set)
    reinit_tasks: Set[str] = field(default_factory=set)
    reinit_tensors: int = 0
    resized: List[str] = field(default_factory=list)
    missing_keys: List[str] = field(default_factory=list)
    unexpected_keys: List[str] = field(default_factory=list)


def _fresh(module_t: Tensors, key: str) -> Optional[torch.Tensor]:
    """The reference's name-based re-initialisation of ``key`` (None when the module has no such tensor)."""
    if key not in module_t:
        return None
    t = torch.empty_like(module_t[key])
    with torch.no_grad():
        if "weight" in key:
            nn.init.xavier_uniform_(t)       # raises ValueError on 1-D tensors, like the reference
        else:
            t.zero_()                        # 'bias' -> zeros; anything else is undefined in the reference
    return t


def _spelled(key: str, ckpt: Tensors) -> Optional[str]:
    """``key`` as the checkpoint spells it: as given, or with the ``model.`` prefix toggled (:62-84)."""
    other = key[6:] if key.startswith("model.") else "model." + key
    return key if key in ckpt else (other if other in ckpt else None)


def interpolate_bias_table(ckpt_tensor: torch.Tensor, model_tensor: torch.Tensor) -> torch.Tensor:
    """Swin relative-position table (N_old, heads) -> (N_new, heads): bicubic, ``align_corners=False``, on the
    square (2w-1) x (2w-1) grid (:33-56).  Non-square lengths raise AssertionError."""
    (n_old, heads), n_new = ckpt_tensor.shape, model_tensor.shape[0]
    if n_old == n_new:
        return ckpt_tensor
    side = [int(n ** 0.5) for n in (n_old, n_new)]
    assert side[0] ** 2 == n_old, f"Checkpoint bias table shape {n_old} is not square"
    assert side[1] ** 2 == n_new, f"Model bias table shape {n_new} is not square"
    grid = ckpt_tensor.reshape(1, side[0], side[0], heads).permute(0, 3, 1, 2)
    grid = torch.nn.functional.interpolate(grid, size=(side[1], side[1]), mode="bicubic", align_corners=False)
    return grid.permute(0, 2, 3, 1).reshape(n_new, heads)


def strip_model_prefix_if_needed(state_dict: Tensors, model_dict: Tensors, verbose: bool = False) -> Tensors:
    if not any(k.startswith("model.") for k in state_dict) or any(k.startswith("model.") for k in model_dict):
        logger.info("→ No prefix stripping needed.")
        return state_dict
    out = {(k[6:] if k.startswith("model.") else k): v for k, v in state_dict.items()}
    logger.info(f"→ Stripped 'model.' prefix from {sum(k.startswith('model.') for k in state_dict)} keys.")
    return out


def _read(path: str) -> Tensors:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return dict(load_file(path))
    blob = torch.load(path, map_location="cpu", weights_only=False)
    return dict(blob.get("state_dict", blob))


def _repair_head(ckpt: Tensors, module_t: Tensors, w_key: str, n_classes: int, label: str, rep: LoadReport) -> bool:
    """One classification head under one spelling of its key.  True when the checkpoint's head fits."""
    b_key = w_key.replace("weight", "bias")
    found_w, found_b = _spelled(w_key, ckpt), _spelled(b_key, ckpt)
    if found_w is not None and ckpt[found_w].shape[0] == n_classes:
        rep.matched_tasks.add(label)
        return True
    if found_w is not None:
        logger.info(f"→ Mismatch: {found_w}: ckpt={ckpt[found_w].shape[0]}, config={n_classes}")
        wanted = [w_key] + ([b_key] if found_b is not None else [])
    else:
        logger.info(f"→ Missing: {w_key}")
        wanted = [w_key, b_key]
    for k in wanted:                      # weight first, then bias: the RNG order of the reference
        t = _fresh(module_t, k)
        if t is not None:
            ckpt[k] = t
            rep.reinit_tensors += 1
    rep.reinit_tasks.add(label)
    return False


def _pass_task_heads(ckpt: Tensors, module_t: Tensors, conf: Dict[str, Any], rep: LoadReport) -> None:
    for task in conf["labels"]:
        n = len(conf["labels_configs"][task]["value_name"])
        spellings = (f"model.main_decoders.{task}.{HEAD}", f"main_decoders.{task}.{HEAD}")
        if not any(_repair_head(ckpt, module_t, k, n, task, rep) for k in spellings):
            logger.info(f"No valid weights found for task '{task}', reinitialized.")


def _pass_aux_heads(ckpt: Tensors, module_t: Tensors, conf: Dict[str, Any], rep: LoadReport) -> None:
    for key in module_t:
        if key.startswith("model.aux_decoders.") and HEAD in key:
            task = key.split(".")[2].split("__")[1]            # model.aux_decoders.<MOD>__<TASK>.… (:59-60)
            n = len(conf["labels_configs"].get(task, {}).get("value_name", []))
            _repair_head(ckpt, module_t, key, n, task, rep)


def _pass_criterion(ckpt: Tensors, module_t: Tensors, conf: Dict[str, Any], rep: LoadReport) -> None:
    for task in conf["labels"]:
        k = f"criterion.{task}.weight"
        if k in ckpt and k in module_t and ckpt[k].shape != module_t[k].shape:
            logger.info(f"→ Reinitializing criterion weights for {task}")
            ckpt[k] = module_t[k].clone()
            rep.reinit_tensors += 1


def _pass_shapes(ckpt: Tensors, module_t: Tensors, conf: Dict[str, Any], rep: LoadReport) -> None:
    for k in list(ckpt):
        if k not in module_t or ckpt[k].shape == module_t[k].shape:
            continue
        if "relative_position_bias_table" in k:
            logger.info(f"→ Interpolating {k}: {tuple(ckpt[k].shape)} → {tuple(module_t[k].shape)}")
            try:
                ckpt[k] = interpolate_bias_table(ckpt[k], module_t[k])
                rep.resized.append(k)
                continue
            except Exception as e:  # noqa: BLE001 -- the reference catches everything here (:268-271)
                logger.info(f"⚠️  Interpolation failed for {k}: {e}. Reinitializing instead.")
        else:
            logger.info(f"→ Shape mismatch for {k}: checkpoint {tuple(ckpt[k].shape)} vs model "
                        f"{tuple(module_t[k].shape)}. Reinitializing...")
        ckpt[k] = _fresh(module_t, k)
        rep.reinit_tensors += 1


REPAIR_PASSES: Tuple[Callable[[Tensors, Tensors, Dict[str, Any], LoadReport], None], ...] = (
    _pass_task_heads, _pass_aux_heads, _pass_criterion, _pass_shapes)


def load_checkpoint(conf: Dict[str, Any], seg_module: nn.Module, exit_on_fail: bool = True) -> None:
    """Same call and effect as the reference's: ``conf['paths']['ckpt_model_path']`` into ``seg_module`` in place.
    The ``LoadReport`` is left on ``seg_module.last_load_report`` (the reference only logs it)."""
    path = conf["paths"]["ckpt_model_path"]
    logger.info(f"→ Loading checkpoint from: {path}")
    if not path or not os.path.isfile(path):
        logger.info("❌ Invalid checkpoint path.")
        if exit_on_fail:
            raise SystemExit()
        return
    ckpt = _read(path)
    logger.info(f"→ Original state dict keys: {len(ckpt)}")
    ckpt = strip_model_prefix_if_needed(ckpt, seg_module.state_dict())
    module_t = seg_module.state_dict()
    rep = LoadReport()
    for repair in REPAIR_PASSES:
        repair(ckpt, module_t, conf, rep)
    result = seg_module.load_state_dict(ckpt, strict=False)
    rep.missing_keys, rep.unexpected_keys = list(result.missing_keys), list(result.unexpected_keys)
    logger.info(f"Checkpoint load summary: matched={sorted(rep.matched_tasks)} reinitialised={sorted(rep.reinit_tasks)} "
                f"tensors_reinit={rep.reinit_tensors} resized={len(rep.resized)} missing={len(rep.missing_keys)} "
                f"unexpected={len(rep.unexpected_keys)}")
    seg_module.last_load_report = rep
    seg_module.last_load_result = result


def save_checkpoint(path: str, seg_module: nn.Module, class_weights: Optional[Dict[str, torch.Tensor]] = None, epoch: int = 0,
                    global_step: int = 0, extra: Optional[Dict[str, Any]] = None) -> str:
    """The other direction of ``load_checkpoint``: the model's current weights in the layout the reference's training run
    leaves on disk and its loader reads back (checkpoint.py:206-212) -- a Lightning ``.ckpt`` whose ``state_dict`` carries the
    LightningModule's names, i.e. ``model.<key>`` for the segmentation model (tasks_module.py:49) and ``criterion.<task>.weight``
    for the loss weights (tasks_module.py:56, module_setup.py:119-200) --, or, for a ``.safetensors`` path, the same tensors
    without the pickle.  With the training engine the parameters are views into the optimizer's arena, so this is the
    trained state.  ``extra``: further top-level entries of the ``.ckpt`` dict (optimizer / scheduler state for a resume)."""
    tensors = {"model." + k: v.detach().to("cpu").contiguous().clone() for k, v in seg_module.state_dict().items()}
    for task, w in (class_weights or {}).items():
        tensors[f"criterion.{task}.weight"] = torch.as_tensor(w, dtype=torch.float32).detach().to("cpu").clone()
    if path.endswith(".safetensors"):
        from safetensors.torch import save_file
        save_file(tensors, path)
        return path
    blob = {"epoch": int(epoch), "global_step": int(global_step), "pytorch-lightning_version": "2.5.1", "state_dict": tensors}
    blob.update(extra or {})
    torch.save(blob, path)
    return path

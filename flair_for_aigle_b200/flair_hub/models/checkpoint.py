"""Checkpoint loading: drop-in for flair_hub/models/checkpoint.py (same names, same behaviour).

Behaviour kept from the reference (checkpoint.py:176-290):
  * ``.safetensors`` via safetensors, anything else via ``torch.load(...)["state_dict"]`` (:206-212);
  * a leading ``model.`` (Lightning) is stripped when the module's own keys have none (:134-173);
  * per task, ``[model.]main_decoders.<task>.seg_model.segmentation_head.0.{weight,bias}`` is
    re-initialised (Xavier / zeros) when missing or when its class count differs (:224-241, :87-131);
  * any other shape mismatch: ``relative_position_bias_table`` is resized bicubically (:33-56),
    everything else re-initialised (:265-274);
  * ``load_state_dict(strict=False)`` (:278); invalid path -> ``SystemExit`` unless
    ``exit_on_fail=False`` (:200-204).
"""
from __future__ import annotations

import logging
import os
from typing import Any, Dict, List, Optional, Set

import torch
import torch.nn as nn

logger = logging.getLogger(__name__)


def reinit_param(state_dict: dict, model_dict: dict, key: str) -> bool:
    if key not in model_dict:
        return False
    with torch.no_grad():
        fresh = torch.empty_like(model_dict[key])
        if 'weight' in key:
            if fresh.dim() >= 2:
                nn.init.xavier_uniform_(fresh)
            else:
                fresh.fill_(1.0)
        elif 'bias' in key:
            fresh.zero_()
        state_dict[key] = fresh
    return True


def interpolate_bias_table(ckpt_tensor: torch.Tensor, model_tensor: torch.Tensor) -> torch.Tensor:
    """(N_old, heads) -> (N_new, heads), bicubic on the square table (checkpoint.py:33-56)."""
    n_old, heads = ckpt_tensor.shape
    n_new = model_tensor.shape[0]
    if n_old == n_new:
        return ckpt_tensor
    s_old, s_new = int(n_old ** 0.5), int(n_new ** 0.5)
    assert s_old * s_old == n_old, f"Checkpoint bias table shape {n_old} is not square"
    assert s_new * s_new == n_new, f"Model bias table shape {n_new} is not square"
    t = ckpt_tensor.reshape(1, s_old, s_old, heads).permute(0, 3, 1, 2)
    t = torch.nn.functional.interpolate(t, size=(s_new, s_new), mode='bicubic', align_corners=False)
    return t.permute(0, 2, 3, 1).reshape(n_new, heads)


def get_task_name_from_aux_key(key: str) -> str:
    return key.split(".")[2].split("__")[1]


def resolve_key(key: str, state_dict: dict) -> Optional[str]:
    alt = key[len("model."):] if key.startswith("model.") else f"model.{key}"
    for k in (key, alt):
        if k in state_dict:
            return k
    return None


def check_and_reinit_layer(state_dict, model_dict, key_weight, key_bias, expected_classes, matched_tasks: Set[str],
                           reinit_tasks: Set[str], task_label: str, reinit_counter: List[int]) -> None:
    kw, kb = resolve_key(key_weight, state_dict), resolve_key(key_bias, state_dict)
    if kw:
        found = state_dict[kw].shape[0]
        if found != expected_classes:
            logger.info(f"→ Mismatch: {kw}: ckpt={found}, config={expected_classes}")
            reinit_counter[0] += reinit_param(state_dict, model_dict, key_weight)
            if kb:
                reinit_counter[0] += reinit_param(state_dict, model_dict, key_bias)
            reinit_tasks.add(task_label)
        else:
            matched_tasks.add(task_label)
    else:
        logger.info(f"→ Missing: {key_weight}")
        if key_weight in model_dict:
            reinit_counter[0] += reinit_param(state_dict, model_dict, key_weight)
        if key_bias in model_dict:
            reinit_counter[0] += reinit_param(state_dict, model_dict, key_bias)
        reinit_tasks.add(task_label)


def strip_model_prefix_if_needed(state_dict: Dict[str, torch.Tensor], model_dict: Dict[str, torch.Tensor],
                                 verbose: bool = False) -> Dict[str, torch.Tensor]:
    ckpt_has = any(k.startswith("model.") for k in state_dict)
    model_has_none = all(not k.startswith("model.") for k in model_dict)
    if not (ckpt_has and model_has_none):
        logger.info("→ No prefix stripping needed.")
        return state_dict
    out, n = {}, 0
    for k, v in state_dict.items():
        if k.startswith("model."):
            out[k[len("model."):]] = v
            n += 1
        else:
            out[k] = v
    logger.info(f"→ Stripped 'model.' prefix from {n} keys.")
    return out


def load_checkpoint(conf: Dict[str, Any], seg_module: nn.Module, exit_on_fail: bool = True) -> None:
    path = conf['paths']['ckpt_model_path']
    logger.info(f"→ Loading checkpoint from: {path}")
    if not path or not os.path.isfile(path):
        logger.info("❌ Invalid checkpoint path.")
        if exit_on_fail:
            raise SystemExit()
        return

    if path.endswith(".safetensors"):
        from safetensors.torch import load_file as safe_load_file
        state_dict = safe_load_file(path)
    else:
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        state_dict = ckpt.get("state_dict", ckpt)
    logger.info(f"→ Original state dict keys: {len(state_dict)}")

    state_dict = strip_model_prefix_if_needed(dict(state_dict), seg_module.state_dict())
    model_dict = seg_module.state_dict()
    tasks = conf["labels"]
    matched, reinit, counter = set(), set(), [0]

    for task in tasks:
        n_classes = len(conf["labels_configs"][task]["value_name"])
        ok = False
        for w_key in (f"model.main_decoders.{task}.seg_model.segmentation_head.0.weight",
                      f"main_decoders.{task}.seg_model.segmentation_head.0.weight"):
            before = len(matched)
            check_and_reinit_layer(state_dict, model_dict, w_key, w_key.replace("weight", "bias"), n_classes, matched,
                                   reinit, task, counter)
            if len(matched) > before:
                ok = True
                break
        if not ok:
            logger.info(f"No valid weights found for task '{task}', reinitialized.")

    for key in model_dict:
        if key.startswith("model.aux_decoders.") and "seg_model.segmentation_head.0.weight" in key:
            task_id = get_task_name_from_aux_key(key)
            n_classes = len(conf["labels_configs"].get(task_id, {}).get("value_name", []))
            check_and_reinit_layer(state_dict, model_dict, key, key.replace("weight", "bias"), n_classes, matched,
                                   reinit, task_id, counter)

    for task in tasks:
        ck = f"criterion.{task}.weight"
        if ck in state_dict and ck in model_dict and state_dict[ck].shape != model_dict[ck].shape:
            state_dict[ck] = model_dict[ck].clone()
            counter[0] += 1

    for k in list(state_dict):
        if k in model_dict and state_dict[k].shape != model_dict[k].shape:
            if "relative_position_bias_table" in k:
                try:
                    state_dict[k] = interpolate_bias_table(state_dict[k], model_dict[k])
                except Exception as e:  # noqa: BLE001 - mirror the reference's catch-all
                    logger.info(f"⚠️  Interpolation failed for {k}: {e}. Reinitializing instead.")
                    counter[0] += reinit_param(state_dict, model_dict, k)
            else:
                logger.info(f"→ Shape mismatch for {k}: checkpoint {tuple(state_dict[k].shape)} vs model "
                            f"{tuple(model_dict[k].shape)}. Reinitializing...")
                counter[0] += reinit_param(state_dict, model_dict, k)

    result = seg_module.load_state_dict(state_dict, strict=False)
    logger.info(f"Checkpoint load summary: matched={sorted(matched)} reinitialised={sorted(reinit)} "
                f"tensors_reinit={counter[0]} missing={len(result.missing_keys)} "
                f"unexpected={len(result.unexpected_keys)}")
    seg_module.last_load_result = result

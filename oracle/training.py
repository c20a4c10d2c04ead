"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU/torch restatement of the training step's loss and optimizer
side, flair_hub/tasks/tasks_module.py:133-167 (``SegmentationTask.step``), flair_hub/tasks/module_setup.py:155,180-196
(``FLAIRLosses``) and tasks_module.py:377-391 (``_init_optimizer``).  The arithmetic is torch's own
(``nn.CrossEntropyLoss``, ``torch.optim.AdamW``), which IS what the reference calls, so this part of the oracle is
pinned by construction; the reference ships no recorded losses to compare with."""
from typing import Dict

import torch
import torch.nn as nn


def default_class_weights(task_config: dict) -> torch.Tensor:
    """module_setup.py:180-196."""
    vw = task_config['value_weights']
    w = torch.FloatTensor([vw['default']] * len(task_config['value_name']))
    if vw.get('default_exceptions'):
        for key, value in vw['default_exceptions'].items():
            w[key] = value
    return w


def step(model: nn.Module, batch: Dict[str, torch.Tensor], config: dict, apply_mod_dropout: bool = False):
    """tasks_module.py:144-167: -> (loss, preds, targets).  The auxiliary loss is left out because it is identically zero in
    the reference (tests/test_reference_pin.py::test_reference_aux_loss_is_identically_zero); ``apply_mod_dropout`` =
    ``self.mod_dropout if training else False`` there (:145)."""
    dict_logits_task, _ = model(batch, apply_mod_dropout) if apply_mod_dropout else model(batch)
    loss_sum = 0
    all_preds, all_targets = {}, {}
    for task, logits in dict_logits_task.items():
        targets = batch[task].to(logits.device)
        targets = torch.argmax(targets, dim=1) if targets.ndim == 4 else targets
        w = default_class_weights(config['labels_configs'][task]).to(logits.device)
        main_loss = nn.CrossEntropyLoss(weight=w)(logits, targets)
        main_preds = torch.argmax(torch.softmax(logits, dim=1), dim=1)
        task_weight = config['labels_configs'][task].get('task_weight', 1.0)
        loss_sum = loss_sum + task_weight * main_loss
        all_preds[task] = main_preds
        all_targets[task] = targets.to(torch.int32)
    return loss_sum, all_preds, all_targets


def init_optimizer(cfg: dict, params):
    """tasks_module.py:377-391."""
    optim_type = cfg['optimizer']
    if optim_type == 'sgd':
        return torch.optim.SGD(params, lr=cfg["learning_rate"])
    if optim_type in ('adam', 'adamw'):
        cls = torch.optim.AdamW if optim_type == 'adamw' else torch.optim.Adam
        return cls(params, lr=cfg["learning_rate"], weight_decay=cfg['optim_weight_decay'], betas=tuple(cfg['optim_betas']))
    raise ValueError(f"Unsupported optimizer type: {optim_type}")

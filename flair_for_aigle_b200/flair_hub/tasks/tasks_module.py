"""``SegmentationTask``: the loss / prediction / optimizer side of flair_hub/tasks/tasks_module.py:133-167,377-391
 on the CUDA kernels of csrc/training_ops.cu.

Built: ``step(batch, training=False)`` -- the reference's validation step: forward (eval mode, the zonal engines),
``targets = argmax(one-hot)``, weighted cross entropy times ``task_weight``, ``preds = argmax(softmax(logits))``;
``loss_gradients()`` -- d loss / d logits per task, the seed of the backward pass; ``AdamW`` -- ``torch.optim.AdamW``'s
update over a flat parameter arena (``_init_optimizer``).  NOT built: the backward of the encoders / decoder (no backward
kernels exist yet), so ``step(batch, training=True)`` raises instead of silently skipping the gradient."""
from typing import Dict, Iterable, List

import torch

from ... import native as nv
from .module_setup import FLAIRLosses


class AdamW:
    """``torch.optim.AdamW(params, lr, weight_decay, betas)`` (tasks_module.py:385-389) as one fused kernel per step:
    parameters, gradients and both moments live in flat fp32 arenas; ``params`` become views into the arena."""

    def __init__(self, params: Iterable[torch.Tensor], lr: float, weight_decay: float = 0.01, betas=(0.9, 0.999),
                 eps: float = 1e-8):
        self.params: List[torch.Tensor] = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise nv.NativeError("AdamW runs on CUDA only (no CPU fallback)")
        self.lr, self.weight_decay, self.betas, self.eps = float(lr), float(weight_decay), tuple(betas), float(eps)
        n = sum(p.numel() for p in self.params)
        self.arena = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)       # the same counter on the device (step_dev())
        self._hyper = torch.zeros(2, dtype=torch.float32, device=dev)
        off = 0
        self.grads: List[torch.Tensor] = []
        for p in self.params:
            k = p.numel()
            self.arena[off:off + k].copy_(p.detach().reshape(-1).float())
            p.data = self.arena[off:off + k].view(p.shape)              # the parameter now aliases the arena
            self.grads.append(self.grad[off:off + k].view(p.shape))
            off += k

    def zero_grad(self) -> None:
        self.grad.zero_()

    def step(self) -> None:
        self.step_count += 1
        self.step_dev.add_(1)
        nv.adamw_step(self.arena, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                      self.weight_decay, self.step_count)

    def step_on_device_counter(self) -> None:
        """The same update driven by the device-resident counter: nothing in the launch depends on the step number, so it can
        be captured in a CUDA graph and replayed (engine/train_step.py).  The caller keeps ``step_count`` in sync."""
        nv.adamw_step_dev(self.arena, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                          self.weight_decay, self.step_dev, self._hyper)


def init_optimizer(cfg: dict, params: Iterable[torch.Tensor]) -> AdamW:
    """tasks_module.py:377-391 (``_init_optimizer``): 'adamw' is built; 'sgd' / 'adam' raise."""
    optim_type = cfg['optimizer']
    if optim_type == 'adamw':
        return AdamW(params, lr=cfg["learning_rate"], weight_decay=cfg['optim_weight_decay'], betas=tuple(cfg['optim_betas']))
    if optim_type in ('sgd', 'adam'):
        raise NotImplementedError(f"optimizer '{optim_type}': only the reference default 'adamw' has a kernel")
    raise ValueError(f"Unsupported optimizer type: {optim_type}")


class SegmentationTask:
    def __init__(self, model, config: dict):
        self.model, self.config = model, config
        self.criterion = FLAIRLosses(config).get_losses()
        self.mod_dropout = False

    def forward(self, batch: Dict[str, torch.Tensor]):
        return self.model(batch)

    @torch.no_grad()
    def step(self, batch: Dict[str, torch.Tensor], training: bool = False):
        """tasks_module.py:133-167.  -> (loss, {task: preds int32 (B,H,W)}, {task: targets int32 (B,H,W)})."""
        if training:
            raise NotImplementedError("step(training=True) returns a loss for an external autograd / optimizer, which does not "
                                      "exist here: use training_step(batch), which runs forward, backward and AdamW on the "
                                      "CUDA kernels (engine/train_step.py)")
        dict_logits_task, _ = self.forward(batch)
        loss_sum = None
        all_preds, all_targets = {}, {}
        for task, logits in dict_logits_task.items():
            targets = batch[task].to(logits.device)
            targets = nv.onehot_argmax(targets) if targets.ndim == 4 else targets.to(torch.int32)
            task_weight = self.config['labels_configs'][task].get('task_weight', 1.0)
            main_loss, preds = self.criterion[task](logits, targets, task_weight=task_weight, want_preds=True)
            if not bool(torch.isfinite(main_loss)):
                raise ValueError(f"Invalid loss for task {task}: {float(main_loss)}")           # tasks_module.py:157
            loss_sum = main_loss if loss_sum is None else loss_sum + main_loss
            all_preds[task], all_targets[task] = preds, targets
        return loss_sum, all_preds, all_targets

    def configure_trainer(self, optim_cfg: dict):
        """``_init_optimizer`` (tasks_module.py:377-391) + the training engine for ``convnextv2_*-unet`` models with one or more
        mono-temporal encoders.  The model's parameters become views into the optimizer's flat arena."""
        from ...engine.convnext_unet import CONVNEXTV2_CFGS
        from ...engine.train_step import ConvNeXtUNetTrainer
        arch = self.config['models']['monotemp_model']['arch'] if 'models' in self.config else self.config['monotemp_arch']
        enc_name, dec_name = arch.rsplit('-', 1)
        enc_name = enc_name[3:] if enc_name.startswith('tu-') else enc_name
        if enc_name not in CONVNEXTV2_CFGS or dec_name.lower() != 'unet':
            raise NotImplementedError(f"training engine: '{arch}' (built for convnextv2_*-unet)")
        if optim_cfg['optimizer'] != 'adamw':
            raise NotImplementedError(f"optimizer '{optim_cfg['optimizer']}': only the reference default 'adamw' has a kernel")
        depths, dims = CONVNEXTV2_CFGS[enc_name]
        task = self.config['labels'][0]
        state = {k: v for k, v in self.model.state_dict(keep_vars=True).items()}
        weight = self.criterion[task].weight
        self.trainer = ConvNeXtUNetTrainer(state, depths, dims,            # the Parameter objects themselves: AdamW rebinds their .data
                                           list(self.model.active_mono), task, weight.to(next(iter(state.values())).device),
                                           task_weight=self.config['labels_configs'][task].get('task_weight', 1.0),
                                           lr=optim_cfg['learning_rate'], weight_decay=optim_cfg['optim_weight_decay'],
                                           betas=tuple(optim_cfg['optim_betas']))
        return self.trainer

    def training_step(self, batch: Dict[str, torch.Tensor]):
        """tasks_module.py:196-207 (training_step -> step(training=True) -> backward -> optimizer.step) in one call.
        -> (loss before the update, {task: preds})."""
        if getattr(self, 'trainer', None) is None:
            raise RuntimeError("call configure_trainer(optim_cfg) first")
        loss, preds = self.trainer.step(batch)
        if hasattr(self.model, '_engines'):
            self.model._engines = {}                 # the inference engines hold repacked copies of the old weights
        return loss, {self.trainer.task: preds}

    def loss_gradients(self) -> Dict[str, torch.Tensor]:
        """d loss_sum / d logits per task for the last ``step`` (fp32, (B,C,H,W))."""
        return {task: crit.backward() for task, crit in self.criterion.items() if crit._saved is not None}

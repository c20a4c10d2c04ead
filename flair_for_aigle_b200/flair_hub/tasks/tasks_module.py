"""``SegmentationTask``: the loss / prediction / optimizer / metric side of flair_hub/tasks/tasks_module.py on
the CUDA kernels of csrc/training_ops.cu and the training engine.

``step(batch, training=False)`` -- the reference's validation step (:133-167): forward (eval mode, the zonal engines),
``targets = argmax(one-hot)``, weighted cross entropy times ``task_weight``, ``preds = argmax(softmax(logits))``;
``configure_trainer`` / ``training_step`` -- forward, backward and AdamW in one call (engine/train_step.py; :196-201,
:377-391); ``validation_step`` and the two epoch-end hooks -- the reference's metric bookkeeping (:63-93, :196-201,
:232-236, :268-276, :302-338: weighted mIoU for training and validation, per-class validation IoU, mean losses) on
tasks/metrics.py, returning the values the reference hands to Lightning's ``self.log`` as a dict.  ``loss_gradients()`` --
d loss / d logits per task.  Modality dropout (``config['modalities']['modality_dropout']``) is applied by the training engine.  ``configure_optimizers``
adds the reference's learning-rate schedules (schedulers.py).  Not built: the auxiliary decoders (their loss is identically zero
in the reference, tests/test_reference_pin.py), the per-class validation loss log."""
from typing import Dict, Iterable, List

import torch

from ... import native as nv
from .module_setup import FLAIRLosses


class StepSegments:
    """Adam step counters of a flat parameter arena.  torch keeps one counter PER PARAMETER and leaves a parameter without a
    gradient untouched (moments, counter and weight decay included); here the arena [0, n) is covered by segments
    [lo, hi, steps] -- one segment until a step skips a range (modality dropout: a whole encoder has no gradient), then the
    segments split at that range's ends and count on their own.  Host bookkeeping only (tests/test_step_segments.py)."""

    def __init__(self, n: int):
        self.items = [[0, int(n), 0]]

    def split(self, at: int) -> None:
        for i, (lo, hi, k) in enumerate(self.items):
            if lo < at < hi:
                self.items[i:i + 1] = [[lo, at, k], [at, hi, k]]
                return

    def advance(self, skip=()):
        """One optimizer step in which the ranges of ``skip`` [(lo, hi), ...] received no gradient: -> [(lo, hi, step)] of the
        segments to update, ``step`` being that segment's own (already incremented) counter."""
        for lo, hi in skip:
            self.split(lo)
            self.split(hi)
        todo = []
        for seg in self.items:
            lo, hi, _ = seg
            if any(a <= lo and hi <= b for a, b in skip):
                continue
            seg[2] += 1
            todo.append((lo, hi, seg[2]))
        return todo

    def __len__(self) -> int:
        return len(self.items)


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
        self._segs = StepSegments(n)                 # per-range step counters (torch: per parameter), see StepSegments
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

    def state_dict(self) -> dict:
        """Everything a resumed run needs besides the weights (those are the model's own state_dict): both moment arenas on
        the CPU, the step counters (per arena segment, like torch's per-parameter ``step``) and the hyper-parameters."""
        return {"exp_avg": self.exp_avg.detach().to("cpu").clone(), "exp_avg_sq": self.exp_avg_sq.detach().to("cpu").clone(),
                "step_count": int(self.step_count), "segments": [list(map(int, seg)) for seg in self._segs.items],
                "lr": self.lr, "weight_decay": self.weight_decay, "betas": tuple(self.betas), "eps": self.eps}

    def load_state_dict(self, state: dict) -> None:
        n = self.arena.numel()
        if state["exp_avg"].numel() != n or state["exp_avg_sq"].numel() != n:
            raise ValueError(f"optimizer state of {state['exp_avg'].numel()} values for an arena of {n}")
        segs = [list(map(int, seg)) for seg in state["segments"]]
        if not segs or segs[0][0] != 0 or segs[-1][1] != n or any(a[1] != b[0] for a, b in zip(segs, segs[1:])):
            raise ValueError("optimizer state: the step-counter segments do not cover the arena")
        self.exp_avg.copy_(state["exp_avg"].to(self.exp_avg.device))
        self.exp_avg_sq.copy_(state["exp_avg_sq"].to(self.exp_avg_sq.device))
        self.step_count = int(state["step_count"])
        self.step_dev.fill_(self.step_count)
        self._segs.items = segs
        self.lr, self.weight_decay = float(state["lr"]), float(state["weight_decay"])
        self.betas, self.eps = tuple(state["betas"]), float(state["eps"])

    def step(self, skip=()) -> None:
        """``skip``: arena ranges [(lo, hi), ...] whose parameters received NO gradient this step: left exactly as they are,
        like ``torch.optim.AdamW`` does for ``p.grad is None`` (no decay, no moment update, their step counter stands still)."""
        self.step_count += 1
        self.step_dev.add_(1)
        for lo, hi, k in self._segs.advance(skip):
            nv.adamw_step(self.arena[lo:hi], self.grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], self.lr,
                          self.betas[0], self.betas[1], self.eps, self.weight_decay, k)

    @property
    def _segments(self):
        return self._segs.items

    @property
    def uniform_steps(self) -> bool:
        """True while every parameter has taken part in every step (the device-counter update assumes it)."""
        return len(self._segs) == 1 and self._segs.items[0][2] == self.step_count

    def step_on_device_counter(self) -> None:
        """The same update driven by the device-resident counter: nothing in the launch depends on the step number, so it can
        be captured in a CUDA graph and replayed (engine/train_step.py).  Host counters are NOT touched here (a capture does
        not execute): the caller reports every executed step with ``note_device_step()``."""
        if len(self._segs) != 1:
            raise RuntimeError("step_on_device_counter(): some parameters skipped earlier steps (per-segment counters)")
        nv.adamw_step_dev(self.arena, self.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps,
                          self.weight_decay, self.step_dev, self._hyper)

    def note_device_step(self) -> None:
        """One step driven by the device counter has run (a graph replay): the host mirrors follow, so that an eager step
        afterwards continues with the right bias corrections."""
        self.step_count += 1
        self._segs.items[0][2] += 1


def init_optimizer(cfg: dict, params: Iterable[torch.Tensor]) -> AdamW:
    """tasks_module.py:377-391 (``_init_optimizer``): 'adamw' is built; 'sgd' / 'adam' raise."""
    optim_type = cfg['optimizer']
    if optim_type == 'adamw':
        return AdamW(params, lr=cfg["learning_rate"], weight_decay=cfg['optim_weight_decay'], betas=tuple(cfg['optim_betas']))
    if optim_type in ('sgd', 'adam'):
        raise NotImplementedError(f"optimizer '{optim_type}': only the reference default 'adamw' has a kernel")
    raise ValueError(f"Unsupported optimizer type: {optim_type}")


class SegmentationTask:
    def __init__(self, model, config: dict, criterion=None):
        """tasks_module.py:37-61.  ``criterion``: ``FLAIRLosses(config).get_losses()`` as module_setup.py:70-77 passes it for
        training; built here when omitted (the reference's predict stage passes none and never needs one)."""
        self.model, self.config = model, config
        if criterion is None and all('value_weights' in config.get('labels_configs', {}).get(t, {}) for t in config.get('labels', [])):
            criterion = FLAIRLosses(config).get_losses()
        self.criterion = criterion                    # None: a predict-stage module (no class weights in its config)
        # tasks_module.py:59-61: the configured probabilities only switch the feature on (the per-step probability is drawn)
        self.mod_dropout = any(v > 0 for v in (self.config.get('modalities', {}).get('modality_dropout', {}) or {}).values())
        self._init_metrics()

    def _init_metrics(self) -> None:
        """tasks_module.py:63-93: weighted mean IoU (train, val), per-class IoU (val), mean loss (train, val), per task."""
        from .metrics import MeanMetric, MulticlassJaccardIndex
        labels = self.config.get('labels', [])
        n_cls = {task: len(self.config['labels_configs'][task]['value_name']) for task in labels}
        self.train_metrics = {task: MulticlassJaccardIndex(n_cls[task], average='weighted') for task in labels}
        self.val_metrics = {task: MulticlassJaccardIndex(n_cls[task], average='weighted') for task in labels}
        self.val_iou = {task: MulticlassJaccardIndex(n_cls[task], average=None) for task in labels}
        self.train_loss, self.val_loss = MeanMetric(), MeanMetric()

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
                                           betas=tuple(optim_cfg['optim_betas']), mod_dropout=self.mod_dropout)
        return self.trainer

    def save_checkpoint(self, path: str, epoch: int = 0) -> str:
        """What Lightning's ModelCheckpoint leaves for this module (``model.*`` / ``criterion.*`` names in ``state_dict``;
        flair_hub/models/checkpoint.py: save_checkpoint), readable by the reference's and this package's ``load_checkpoint``;
        with a trainer the ``.ckpt`` also carries the optimizer's moments / step counters and the schedule position, which
        ``load_training_state`` puts back for a resume."""
        from ..models.checkpoint import save_checkpoint
        weights = {t: c.weight for t, c in (self.criterion or {}).items() if getattr(c, 'weight', None) is not None}
        extra = {}
        tr = getattr(self, 'trainer', None)
        if tr is not None and not path.endswith(".safetensors"):
            extra["optimizer_states"] = [tr.opt.state_dict()]
            extra["scheduler_position"] = {"global_step": int(getattr(self, '_global_step', 0)),
                                           "using_plateau": bool(getattr(self, '_using_plateau', False))}
        return save_checkpoint(path, self.model, weights, epoch=epoch, global_step=int(getattr(self, '_global_step', 0)),
                               extra=extra)

    def load_training_state(self, path: str) -> None:
        """Resume: the optimizer state written by ``save_checkpoint`` back into the trainer (the weights are loaded the usual
        way, ``load_checkpoint``, BEFORE ``configure_trainer``).  Captured CUDA graphs are dropped."""
        blob = torch.load(path, map_location="cpu", weights_only=False)
        tr = getattr(self, 'trainer', None)
        if tr is None:
            raise RuntimeError("load_training_state: call configure_trainer / configure_optimizers first")
        if not blob.get("optimizer_states"):
            raise ValueError(f"{path}: no optimizer state inside")
        tr.opt.load_state_dict(blob["optimizer_states"][0])
        tr.set_lr(tr.opt.lr)
        tr._graph, tr._segments, tr._graph_out = None, [], None
        pos = blob.get("scheduler_position", {})
        self._global_step, self._using_plateau = int(pos.get("global_step", 0)), bool(pos.get("using_plateau", False))

    def configure_optimizers(self, total_steps: int):
        """tasks_module.py:344-376 without Lightning: builds the trainer from ``config['hyperparams']`` (``configure_trainer``)
        and the schedule it names -- 'reduce_on_plateau' (factor 0.5, ``plateau_patience``, cooldown 4, floor 1e-7; stepped by
        ``on_validation_epoch_end`` with the validation loss), 'one_cycle_lr' (peak = ``learning_rate``, ``warmup_fraction`` of
        ``total_steps`` rising from lr/1000; stepped after every training batch), 'cycle_then_plateau' (a one-cycle rise over
        ``warmup_fraction * total_steps`` steps, then plateau halving with patience 10; :213-231, :311-314) or none.
        ``total_steps`` is Lightning's ``trainer.estimated_stepping_batches``.  Returns what the reference returns, with this
        package's optimizer / scheduler objects (flair_hub/tasks/schedulers.py)."""
        from .schedulers import OneCycleLR, ReduceLROnPlateau
        cfg = self.config['hyperparams']
        if getattr(self, 'trainer', None) is None:
            self.configure_trainer(cfg)
        tr = self.trainer
        self._scheduler_type = cfg.get("scheduler", None)
        self._scheduler = self._warmup_scheduler = self._plateau_scheduler = None
        self._using_plateau, self._global_step = False, 0
        warmup_fraction = cfg.get("warmup_fraction", 0.0)
        if self._scheduler_type == "reduce_on_plateau":
            self._scheduler = ReduceLROnPlateau(tr.set_lr, tr.opt.lr, factor=0.5, patience=cfg['plateau_patience'], cooldown=4,
                                                min_lr=1e-7)
            return {"optimizer": tr.opt, "lr_scheduler": {"scheduler": self._scheduler, "monitor": "val_loss", "interval": "epoch"}}
        if self._scheduler_type == "one_cycle_lr":
            tr.cuda_graph = False                    # the rate changes every step: a captured graph would freeze it
            self._scheduler = OneCycleLR(tr.set_lr, max_lr=cfg["learning_rate"], total_steps=total_steps,
                                         pct_start=warmup_fraction, div_factor=1000)
            return {"optimizer": tr.opt, "lr_scheduler": {"scheduler": self._scheduler, "interval": "step"}}
        if self._scheduler_type == "cycle_then_plateau":
            warmup_steps = int(warmup_fraction * total_steps)
            self._warmup_scheduler = OneCycleLR(tr.set_lr, max_lr=cfg["learning_rate"], total_steps=warmup_steps, pct_start=1.0,
                                                div_factor=1000, final_div_factor=1)
            self._plateau_scheduler = ReduceLROnPlateau(tr.set_lr, cfg["learning_rate"], factor=0.5, patience=10, cooldown=4,
                                                        min_lr=1e-7)
            return {"optimizer": tr.opt}
        return tr.opt

    def _after_train_batch(self) -> None:
        """What Lightning (interval 'step') and the reference's ``on_train_batch_end`` (:213-231) do after an optimizer step."""
        self._global_step = getattr(self, '_global_step', 0) + 1
        kind = getattr(self, '_scheduler_type', None)
        if kind == "one_cycle_lr" and self._global_step < self._scheduler.total_steps:
            self._scheduler.step()
        elif kind == "cycle_then_plateau" and not self._using_plateau:
            if self._global_step < self._warmup_scheduler.total_steps:
                self._warmup_scheduler.step()
            if self._global_step == self._warmup_scheduler.total_steps:
                self._using_plateau = True
                self._plateau_scheduler.lr = self.trainer.opt.lr      # the plateau phase starts from the rate the warm-up reached

    def training_step(self, batch: Dict[str, torch.Tensor]):
        """tasks_module.py:196-207 (training_step -> step(training=True) -> backward -> optimizer.step) in one call.
        -> (loss before the update, {task: preds})."""
        if getattr(self, 'trainer', None) is None:
            raise RuntimeError("call configure_trainer(optim_cfg) first")
        loss, preds = self.trainer.step(batch)
        if hasattr(self.model, '_engines'):
            self.model._engines = {}                 # the inference engines hold repacked copies of the old weights
        task = self.trainer.task
        self.train_loss.update(loss)                                                  # tasks_module.py:198
        if task in self.train_metrics and self.trainer.last_targets is not None:
            self.train_metrics[task].update(preds, self.trainer.last_targets)         # :199-200
        self._after_train_batch()
        return loss, {task: preds}

    @torch.no_grad()
    def predict_step(self, batch: Dict[str, torch.Tensor], batch_idx: int = 0, dataloader_idx: int = 0):
        """tasks_module.py:337-342: ``{"preds_<task>": argmax over the classes}`` per task, int64 (B,H,W).  The softmax in
        front of the reference's argmax is monotonic and is skipped; the argmax is the ``convert`` kernel (first maximum
        wins, like torch.argmax), one sample at a time."""
        dict_logits_task, _ = self.model(batch)
        return {f"preds_{task}": torch.stack([nv.convert(sample.contiguous(), 0)[0] for sample in logits]).long()
                for task, logits in dict_logits_task.items()}

    def validation_step(self, batch: Dict[str, torch.Tensor]):
        """tasks_module.py:268-276 without the per-class loss log: -> the step's loss."""
        loss, all_preds, all_targets = self.step(batch, training=False)
        self.val_loss.update(loss)
        for task in all_preds:
            self.val_metrics[task].update(all_preds[task], all_targets[task])
            self.val_iou[task].update(all_preds[task], all_targets[task])
        return loss

    def on_train_epoch_end(self) -> Dict[str, float]:
        """tasks_module.py:232-236,258-267: -> {'train_miou_<task suffix>': ..., 'train_loss': ...}; metrics are reset."""
        out = {}
        for task, metric in self.train_metrics.items():
            if metric.confmat is not None:
                out[f"train_miou_{task.split('-')[-1]}"] = float(metric.compute())
            metric.reset()
        if self.train_loss.total is not None:
            out["train_loss"] = float(self.train_loss.compute())
        self.train_loss.reset()
        return out

    def on_validation_epoch_end(self) -> Dict[str, float]:
        """tasks_module.py:302-338: val_loss, val_miou_<task>, val_iou_<task>_<k>_<class name>, val_miou (mean over tasks)."""
        out = {"val_loss": float(self.val_loss.compute())}
        total = 0.0
        for task in self.val_metrics:
            suffix = task.split('-')[-1]
            miou = float(self.val_metrics[task].compute())
            total += miou
            out[f"val_miou_{suffix}"] = miou
            names = self.config['labels_configs'][task]['value_name']
            per_class = torch.nan_to_num(self.val_iou[task].compute(), nan=0.0).tolist()
            for k, iou in enumerate(per_class):
                name = names.get(k, f"class_{k}") if isinstance(names, dict) else f"class_{k}"
                out[f"val_iou_{suffix}_{k}_{name}"] = iou
            self.val_metrics[task].reset()
            self.val_iou[task].reset()
        out["val_miou"] = total / max(1, len(self.val_metrics))
        self.val_loss.reset()
        kind = getattr(self, '_scheduler_type', None)                      # :311-314, and Lightning's epoch-interval monitor
        if kind == "reduce_on_plateau":
            self._scheduler.step(out["val_loss"])
        elif kind == "cycle_then_plateau" and self._using_plateau:
            self._plateau_scheduler.step(out["val_loss"])
        return out

    def loss_gradients(self) -> Dict[str, torch.Tensor]:
        """d loss_sum / d logits per task for the last ``step`` (fp32, (B,C,H,W))."""
        return {task: crit.backward() for task, crit in self.criterion.items() if crit._saved is not None}

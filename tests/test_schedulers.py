"""The restated learning-rate schedules (flair_hub/tasks/schedulers.py) against torch's own scheduler classes, with the
parameterisations the reference's ``configure_optimizers`` uses (tasks_module.py:351-373)."""
import numpy as np
import pytest
import torch

from flair_for_aigle_b200.flair_hub.tasks.schedulers import OneCycleLR, ReduceLROnPlateau


class _Holder:
    def __init__(self, lr):
        self.lr = lr

    def set_lr(self, lr):
        self.lr = lr


def _torch_opt(lr):
    return torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=lr)


@pytest.mark.parametrize("total,pct,kw", [(200, 0.1, {"div_factor": 1000}), (37, 0.3, {"div_factor": 1000}), (1000, 0.0, {"div_factor": 1000}),
                                          (25, 1.0, {"div_factor": 1000, "final_div_factor": 1}),      # the warm-up of cycle_then_plateau
                                          (50, 0.25, {})])
def test_one_cycle_equals_torch(total, pct, kw):
    max_lr = 5e-5
    opt = _torch_opt(max_lr)
    ref = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=max_lr, total_steps=total, pct_start=pct, cycle_momentum=False, **kw)
    h = _Holder(max_lr)
    mine = OneCycleLR(h.set_lr, max_lr=max_lr, total_steps=total, pct_start=pct, **kw)
    for step in range(total):
        assert h.lr == opt.param_groups[0]["lr"], (step, h.lr, opt.param_groups[0]["lr"])
        assert mine.get_last_lr() == ref.get_last_lr()
        opt.step()
        if step < total - 1:
            ref.step()
            mine.step()
    # one step too many extrapolates (like torch), two raise -- with pct_start = 1 the extrapolation itself divides by zero,
    # in torch as well
    with pytest.raises((ValueError, ZeroDivisionError)):
        mine.step()
        mine.step()
    with pytest.raises((ValueError, ZeroDivisionError)):
        ref.step()
        ref.step()


@pytest.mark.parametrize("factor,patience,cooldown,min_lr", [(0.5, 3, 4, 1e-7), (0.5, 10, 4, 1e-7), (0.5, 0, 0, 1e-7), (0.1, 2, 1, 1e-5)])
def test_reduce_on_plateau_equals_torch(factor, patience, cooldown, min_lr):
    rng = np.random.default_rng(patience * 10 + cooldown)
    for trial in range(5):
        lr0 = 2e-4
        opt = _torch_opt(lr0)
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=factor, patience=patience, cooldown=cooldown,
                                                         min_lr=min_lr)
        h = _Holder(lr0)
        mine = ReduceLROnPlateau(h.set_lr, lr0, factor=factor, patience=patience, cooldown=cooldown, min_lr=min_lr)
        loss = 3.0
        for epoch in range(120):
            loss = loss * (1.0 - 0.02 * rng.random()) if rng.random() < 0.35 else loss * (1.0 + 0.01 * rng.random())
            ref.step(loss)
            mine.step(torch.tensor(loss) if epoch % 2 else loss)
            assert h.lr == opt.param_groups[0]["lr"], (trial, epoch)
        assert h.lr < lr0                       # the sequence did plateau


class _FakeOpt:
    def __init__(self, lr):
        self.lr = lr


class _FakeTrainer:
    """Stands in for engine.train_step.ConvNeXtUNetTrainer (CUDA only): what the task module's schedule plumbing touches."""

    def __init__(self, lr):
        self.opt, self.task, self.last_targets, self.cuda_graph = _FakeOpt(lr), "AERIAL_LABEL-COSIA", None, True

    def set_lr(self, lr):
        self.opt.lr = float(lr)

    def step(self, batch):
        return torch.tensor(1.0), torch.zeros(1, 2, 2, dtype=torch.int32)


def _task(hyper):
    from flair_for_aigle_b200.flair_hub.tasks.tasks_module import SegmentationTask
    task_name = "AERIAL_LABEL-COSIA"
    cfg = {"labels": [task_name], "hyperparams": hyper,
           "labels_configs": {task_name: {"value_name": {k: f"c{k}" for k in range(19)}, "task_weight": 1.0,
                                          "value_weights": {"default": 1, "default_exceptions": {}}}},
           "modalities": {"inputs": {"AERIAL_RGBI": True}, "aux_loss": {}, "modality_dropout": {}}}
    task = SegmentationTask(model=None, config=cfg)
    task.trainer = _FakeTrainer(hyper["learning_rate"])
    return task


def test_task_one_cycle_follows_lightning_step_interval():
    """'one_cycle_lr' (tasks_module.py:357-362): Lightning steps the scheduler after every optimizer step."""
    hyper = {"optimizer": "adamw", "learning_rate": 5e-5, "optim_weight_decay": 0.01, "optim_betas": [0.9, 0.999],
             "scheduler": "one_cycle_lr", "warmup_fraction": 0.2}
    total = 40
    task = _task(hyper)
    ret = task.configure_optimizers(total)
    assert ret["lr_scheduler"]["interval"] == "step" and task.trainer.cuda_graph is False
    opt = _torch_opt(hyper["learning_rate"])
    ref = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=hyper["learning_rate"], total_steps=total, pct_start=0.2,
                                              cycle_momentum=False, div_factor=1000)
    for step in range(total):
        assert task.trainer.opt.lr == opt.param_groups[0]["lr"], step
        task.training_step({})
        opt.step()
        if step < total - 1:
            ref.step()


def test_task_cycle_then_plateau_follows_the_reference_hooks():
    """'cycle_then_plateau' (tasks_module.py:364-373 + on_train_batch_end :223-231 + on_validation_epoch_end :311-314), replayed
    here with torch's schedulers and the reference's conditions, against the task module's own plumbing."""
    hyper = {"optimizer": "adamw", "learning_rate": 2e-4, "optim_weight_decay": 0.01, "optim_betas": [0.9, 0.999],
             "scheduler": "cycle_then_plateau", "warmup_fraction": 0.1}
    total, per_epoch = 300, 10
    task = _task(hyper)
    assert not isinstance(task.configure_optimizers(total), dict) or "lr_scheduler" not in task.configure_optimizers(total)
    opt = _torch_opt(hyper["learning_rate"])
    warm = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=hyper["learning_rate"], total_steps=int(0.1 * total), pct_start=1.0,
                                               cycle_momentum=False, div_factor=1000, final_div_factor=1)
    plateau = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.5, patience=10, cooldown=4, min_lr=1e-7)
    using_plateau, global_step = False, 0
    rng = np.random.default_rng(1)
    val = 2.0
    for epoch in range(total // per_epoch):
        for _ in range(per_epoch):
            assert task.trainer.opt.lr == opt.param_groups[0]["lr"], (epoch, global_step)
            task.training_step({})
            opt.step()
            global_step += 1
            if not using_plateau:                                  # the reference's on_train_batch_end
                if global_step < warm.total_steps:
                    warm.step()
                if global_step == warm.total_steps:
                    using_plateau = True
        val = val * (0.97 if epoch < 6 else 1.001 + 0.002 * rng.random())      # improves, then stalls
        task.val_loss.update(val)
        for m in list(task.val_metrics.values()) + list(task.val_iou.values()):   # epoch end needs a confusion matrix
            m.confmat = torch.eye(19, dtype=torch.int64)
        out = task.on_validation_epoch_end()
        assert abs(out["val_loss"] - val) < 1e-6
        if using_plateau:
            plateau.step(out["val_loss"])
    assert task.trainer.opt.lr == opt.param_groups[0]["lr"] < hyper["learning_rate"]

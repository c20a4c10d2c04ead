"""The learning-rate schedules the reference's ``configure_optimizers`` builds (flair_hub/tasks/tasks_module.py:344-376) from
``torch.optim.lr_scheduler``: ``OneCycleLR`` (cosine annealing, two phases, no momentum cycling) and ``ReduceLROnPlateau``
(mode 'min', relative threshold 1e-4).  torch's classes insist on a ``torch.optim.Optimizer``; the training engine's AdamW is a
fused kernel over a flat arena, so the same host arithmetic is restated here (torch 2.x semantics; pinned against torch's own
classes value for value in tests/test_schedulers.py) and writes the rate through ``set_lr`` (any callable, e.g.
``ConvNeXtUNetTrainer.set_lr``).  Host floats only -- nothing here touches the GPU."""
import math
from typing import Callable


class OneCycleLR:
    """``torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr, total_steps, pct_start, anneal_strategy='cos',
    cycle_momentum=False, div_factor, final_div_factor, three_phase=False)``: the rate at construction is
    ``max_lr / div_factor``; every ``step()`` moves one step along  initial -> max_lr  (until ``pct_start * total_steps - 1``)
    -> ``initial / final_div_factor``  (until ``total_steps - 1``), each leg a half cosine."""

    def __init__(self, set_lr: Callable[[float], None], max_lr: float, total_steps: int, pct_start: float = 0.3,
                 div_factor: float = 25.0, final_div_factor: float = 1e4):
        if total_steps <= 0:
            raise ValueError(f"Expected positive integer total_steps, but got {total_steps}")
        if pct_start < 0 or pct_start > 1:
            raise ValueError(f"Expected float between 0 and 1 pct_start, but got {pct_start}")
        self.set_lr, self.total_steps = set_lr, int(total_steps)
        self.max_lr = float(max_lr)
        self.initial_lr = self.max_lr / div_factor
        self.min_lr = self.initial_lr / final_div_factor
        self.phases = [(float(pct_start * self.total_steps) - 1, self.initial_lr, self.max_lr),
                       (self.total_steps - 1, self.max_lr, self.min_lr)]
        self.last_epoch = 0
        self._last_lr = self._lr_at(0)
        self.set_lr(self._last_lr)

    @staticmethod
    def _cos(start: float, end: float, pct: float) -> float:
        return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1)

    def _lr_at(self, step_num: int) -> float:
        if step_num > self.total_steps:
            raise ValueError(f"Tried to step {step_num} times. The specified number of total steps is {self.total_steps}")
        start_step = 0.0
        for i, (end_step, a, b) in enumerate(self.phases):
            if step_num <= end_step or i == len(self.phases) - 1:
                return self._cos(a, b, (step_num - start_step) / (end_step - start_step))
            start_step = end_step
        raise AssertionError("unreachable")

    def step(self) -> None:
        self.last_epoch += 1
        self._last_lr = self._lr_at(self.last_epoch)
        self.set_lr(self._last_lr)

    def get_last_lr(self):
        return [self._last_lr]


class ReduceLROnPlateau:
    """``torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor, patience, threshold=1e-4,
    threshold_mode='rel', cooldown, min_lr, eps=1e-8)``: ``step(metric)`` once per validation epoch."""

    def __init__(self, set_lr: Callable[[float], None], lr: float, factor: float = 0.1, patience: int = 10, cooldown: int = 0,
                 min_lr: float = 0.0, threshold: float = 1e-4, eps: float = 1e-8):
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        self.set_lr, self.lr = set_lr, float(lr)
        self.factor, self.patience, self.cooldown, self.min_lr = factor, patience, cooldown, min_lr
        self.threshold, self.eps = threshold, eps
        self.best, self.num_bad_epochs, self.cooldown_counter, self.last_epoch = math.inf, 0, 0, 0

    def step(self, metric) -> None:
        current = float(metric)
        self.last_epoch += 1
        if current < self.best * (1.0 - self.threshold):
            self.best, self.num_bad_epochs = current, 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            new_lr = max(self.lr * self.factor, self.min_lr)
            if self.lr - new_lr > self.eps:
                self.lr = new_lr
                self.set_lr(new_lr)
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0

    def get_last_lr(self):
        return [self.lr]

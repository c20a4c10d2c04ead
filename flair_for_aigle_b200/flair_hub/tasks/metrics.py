"""The two torchmetrics classes the reference's task module uses (tasks_module.py:6-7,63-93), with the confusion matrix counted
by a CUDA kernel (csrc/training_ops.cu ``confusion_kernel``) instead of ``bincount(target * C + preds)``.

``MulticlassJaccardIndex(num_classes, average)``, ``average`` in {'weighted', 'macro', 'micro', 'none', None}: state = the int64
[C][C] confusion matrix summed over ``update(preds, target)`` calls (rows = target, columns = prediction); ``compute()`` follows
torchmetrics 1.7.0 (the reference's pin, requirements.txt:10) ``_jaccard_index_reduce``: the matrix is cast to float32, IoU_c =
TP / (row sum + column sum - TP) with 0 where a class has no pixels at all, 'weighted' averages with the label counts (row
sums), 'macro' over the classes that occur in labels or predictions.  ``MeanMetric``: running mean of the values it is given."""
from typing import Optional

import torch

from ... import native as nv


class MulticlassJaccardIndex:
    def __init__(self, num_classes: int, average: Optional[str] = "macro"):
        if average not in ("weighted", "macro", "micro", "none", None):
            raise ValueError(f"average={average!r}: expected 'weighted', 'macro', 'micro', 'none' or None")
        self.num_classes, self.average = int(num_classes), average
        self.confmat: Optional[torch.Tensor] = None

    def to(self, device):                      # the reference moves its metric modules explicitly (tasks_module.py:113-120)
        if self.confmat is not None:
            self.confmat = self.confmat.to(device)
        return self

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        """preds / target: integer class maps of one shape (what ``SegmentationTask.step`` returns)."""
        if preds.device.type != "cuda":
            raise nv.NativeError("MulticlassJaccardIndex counts on CUDA only (no CPU fallback)")
        if self.confmat is None:
            self.confmat = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64, device=preds.device)
        nv.confusion_matrix(target, preds, self.num_classes, out=self.confmat)

    def reset(self) -> None:
        self.confmat = None

    def compute(self) -> torch.Tensor:
        if self.confmat is None:
            raise RuntimeError("compute() before update()")
        return jaccard_from_confmat(self.confmat, self.average)


def jaccard_from_confmat(confmat: torch.Tensor, average: Optional[str]) -> torch.Tensor:
    cm = confmat.float()
    num = torch.diag(cm)
    denom = cm.sum(0) + cm.sum(1) - num
    if average == "micro":
        num, denom = num.sum(), denom.sum()
    jaccard = torch.where(denom != 0, num / torch.where(denom != 0, denom, torch.ones_like(denom)), torch.zeros_like(num))
    if average in (None, "none", "micro"):
        return jaccard
    if average == "weighted":
        weights = cm.sum(1)
    else:
        weights = torch.ones_like(jaccard)
        weights[cm.sum(1) + cm.sum(0) == 0] = 0.0
    return ((weights * jaccard) / weights.sum()).sum()


class MeanMetric:
    """torchmetrics.aggregation.MeanMetric for scalar updates: sum of values / number of updates (float32 like there)."""

    def __init__(self):
        self.reset()

    def update(self, value) -> None:
        v = value.detach().float().reshape(-1) if torch.is_tensor(value) else torch.tensor([float(value)])
        self.total = v.sum() if self.total is None else self.total + v.sum().to(self.total.device)
        self.weight += v.numel()

    def compute(self) -> torch.Tensor:
        if self.total is None:
            raise RuntimeError("compute() before update()")
        return self.total / self.weight

    def reset(self) -> None:
        self.total, self.weight = None, 0

    def to(self, device):
        if self.total is not None:
            self.total = self.total.to(device)
        return self

"""CPU restatement of the metric side of the training / evaluation loop.  TEST INFRASTRUCTURE: only tests/, smoke() and the
bench's cpu_baseline leg may import this package.

* ``confusion_matrix``: what both of the reference's call sites compute -- sklearn ``confusion_matrix(target, pred,
  labels=range(C))`` (flair_hub/writer/prediction_writer.py:64,149) and torchmetrics' multiclass update
  (``bincount(target * C + preds, minlength=C*C).reshape(C, C)``): rows = labels, columns = predictions, values outside
  0..C-1 dropped.  Pinned against sklearn 1.x (installed here; the reference pins 1.6.1) in tests/test_metrics.py.
* ``jaccard``: torchmetrics 1.7.0 (requirements.txt:10) ``_jaccard_index_reduce`` for the averages the reference asks for
  (tasks_module.py:74-90: 'weighted' for train / val mIoU, None for the per-class IoU): float32 arithmetic on the matrix.
  torchmetrics is not installed in this image: PARITY UNPINNED for this function (restated from the published source);
  the identities weighted = sum_c support_c IoU_c / sum_c support_c and IoU_c = TP / (TP + FP + FN) are what the tests check.
* the percent scores of flair_hub/writer/metrics_core.py:4-49 need no restatement here: that file is plain numpy, the pin test
  imports it from /root/reference and compares the product's mirror with it bit for bit."""
import numpy as np


def confusion_matrix(target: np.ndarray, pred: np.ndarray, num_classes: int) -> np.ndarray:
    t = np.asarray(target).reshape(-1).astype(np.int64)
    p = np.asarray(pred).reshape(-1).astype(np.int64)
    keep = (t >= 0) & (t < num_classes) & (p >= 0) & (p < num_classes)
    return np.bincount(t[keep] * num_classes + p[keep], minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def jaccard(confmat: np.ndarray, average):
    cm = np.asarray(confmat).astype(np.float32)
    num = np.diag(cm).copy()
    denom = cm.sum(0) + cm.sum(1) - num
    if average == "micro":
        num, denom = num.sum(dtype=np.float32), denom.sum(dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = np.where(denom != 0, num / np.where(denom != 0, denom, np.float32(1)), np.float32(0)).astype(np.float32)
    if average in (None, "none", "micro"):
        return iou
    if average == "weighted":
        w = cm.sum(1)
    else:
        w = np.ones_like(iou)
        w[cm.sum(1) + cm.sum(0) == 0] = 0
    return ((w * iou) / w.sum(dtype=np.float32)).sum(dtype=np.float32)

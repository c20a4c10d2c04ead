"""Scores derived from a confusion matrix (rows = labels, columns = predictions), in percent, with the reference's names and
return conventions (flair_hub/writer/metrics_core.py:4-49).  Host arithmetic on a C x C matrix; the matrix itself is counted
on the GPU (``native.confusion_matrix`` / ``tasks.metrics``).  A class that never occurs scores 0, not NaN."""
import numpy as np


def _percent(num: np.ndarray, den: np.ndarray) -> np.ndarray:
    """100 * num / den with 0 where the ratio is undefined (0 / 0), as the reference's nan -> 0 replacement does."""
    with np.errstate(divide="ignore", invalid="ignore"):
        out = 100 * np.asarray(num, dtype=np.float64) / np.asarray(den, dtype=np.float64)
    out[np.isnan(out)] = 0
    return out


def overall_accuracy(npcm: np.ndarray) -> float:
    """metrics_core.py:4-9."""
    return 100 * (np.trace(npcm) / npcm.sum())


def class_IoU(npcm: np.ndarray, n_class: int) -> tuple:
    """metrics_core.py:12-18 -> (per-class IoU, their mean).  ``n_class`` is accepted and unused, like there."""
    hit = np.diag(npcm)
    ious = _percent(hit, npcm.sum(axis=1) + npcm.sum(axis=0) - hit)
    return ious, np.mean(ious)


def class_precision(npcm: np.ndarray) -> tuple:
    """metrics_core.py:21-27."""
    precision = _percent(np.diag(npcm), npcm.sum(axis=0))
    return precision, np.mean(precision)


def class_recall(npcm: np.ndarray) -> tuple:
    """metrics_core.py:30-36."""
    recall = _percent(np.diag(npcm), npcm.sum(axis=1))
    return recall, np.mean(recall)


def class_fscore(precision: np.ndarray, recall: np.ndarray) -> tuple:
    """metrics_core.py:39-45: harmonic mean of the two percent vectors."""
    with np.errstate(divide="ignore", invalid="ignore"):
        fscore = 2 * (precision * recall) / (precision + recall)
    fscore[np.isnan(fscore)] = 0
    return fscore, np.mean(fscore)

"""Metric side of the training / evaluation loop: the oracle's confusion matrix against sklearn (what
prediction_writer.py:64 calls), the product's ``metrics_core`` mirror against the REFERENCE's own file, and the restated
torchmetrics Jaccard reduce against its defining identities."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import metrics as om
from flair_for_aigle_b200.flair_hub.writer import metrics_core as mc
from flair_for_aigle_b200.flair_hub.tasks.metrics import MeanMetric, jaccard_from_confmat

REF = "/root/reference/flair_hub/writer/metrics_core.py"


def _random_labels(rng, n, C, junk=True):
    t = rng.integers(0, C, n)
    p = np.where(rng.random(n) < 0.7, t, rng.integers(0, C, n))
    if junk:
        t = np.where(rng.random(n) < 0.05, rng.choice([-100, -1, C, 255]), t)
    return t, p


@pytest.mark.parametrize("C,n", [(19, 50_000), (2, 1000), (7, 1), (19, 0)])
def test_oracle_confusion_matrix_equals_sklearn(C, n):
    sk = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(C * 1000 + n)
    t, p = _random_labels(rng, n, C)
    got = om.confusion_matrix(t, p, C)
    assert got.shape == (C, C) and got.dtype == np.int64
    if n:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = sk.confusion_matrix(t, p, labels=list(range(C)))
        assert np.array_equal(got, ref)
    else:
        assert got.sum() == 0


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree is not on this machine")
def test_metrics_core_mirror_equals_the_reference_file():
    spec = importlib.util.spec_from_file_location("ref_metrics_core", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(3)
    for case in range(20):
        C = int(rng.integers(2, 20))
        cm = rng.integers(0, 10_000, (C, C)).astype(np.int64)
        if case % 3 == 0:                      # a class that never occurs, as label or as prediction
            k = int(rng.integers(0, C))
            cm[k, :] = 0
            cm[:, k] = 0
        if case % 5 == 0:
            cm = cm / cm.sum()                 # the "normalised" matrix the reference's writer passes in
        with np.errstate(divide="ignore", invalid="ignore"):
            r_iou, r_miou = ref.class_IoU(cm, C)
            r_p, r_mp = ref.class_precision(cm)
            r_r, r_mr = ref.class_recall(cm)
            r_f, r_mf = ref.class_fscore(r_p, r_r)
            r_oa = ref.overall_accuracy(cm)
        iou, miou = mc.class_IoU(cm, C)
        p, mp = mc.class_precision(cm)
        r, mr = mc.class_recall(cm)
        f, mf = mc.class_fscore(p, r)
        for a, b in ((iou, r_iou), (p, r_p), (r, r_r), (f, r_f)):
            assert np.array_equal(a, b)                          # bit for bit
        assert (miou, mp, mr, mf, mc.overall_accuracy(cm)) == (r_miou, r_mp, r_mr, r_mf, r_oa)


def test_jaccard_reduce_identities_and_product_equals_oracle():
    rng = np.random.default_rng(11)
    for case in range(10):
        C = int(rng.integers(2, 20))
        cm = rng.integers(0, 100_000, (C, C)).astype(np.int64)
        if case % 2 == 0:
            k = int(rng.integers(0, C))
            cm[k, :] = 0
            cm[:, k] = 0
        tp = np.diag(cm).astype(np.float64)
        fp, fn = cm.sum(0) - tp, cm.sum(1) - tp
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = np.where(tp + fp + fn > 0, tp / (tp + fp + fn), 0.0)
        per_class = om.jaccard(cm, None)
        assert per_class.dtype == np.float32 and np.allclose(per_class, iou, rtol=1e-6, atol=0)
        support = cm.sum(1).astype(np.float64)
        assert np.isclose(om.jaccard(cm, "weighted"), (support * iou).sum() / support.sum(), rtol=1e-6)
        present = (cm.sum(1) + cm.sum(0)) > 0
        assert np.isclose(om.jaccard(cm, "macro"), iou[present].mean(), rtol=1e-6)
        assert np.isclose(om.jaccard(cm, "micro"), tp.sum() / (tp + fp + fn).sum(), rtol=1e-6)
        for avg in (None, "none", "weighted", "macro", "micro"):
            got = jaccard_from_confmat(torch.from_numpy(cm), avg).numpy()
            if avg in (None, "none"):
                assert np.array_equal(got, om.jaccard(cm, avg)), avg         # exact sums (< 2^24), one float32 division
            else:                                                            # float32 reductions: summation order differs
                assert np.isclose(got, om.jaccard(cm, avg), rtol=1e-6, atol=0), avg


def test_mean_metric():
    m = MeanMetric()
    with pytest.raises(RuntimeError):
        m.compute()
    for v in (1.0, torch.tensor(2.0), torch.tensor([3.0])):
        m.update(v)
    assert float(m.compute()) == 2.0
    m.reset()
    m.update(5.0)
    assert float(m.compute()) == 5.0

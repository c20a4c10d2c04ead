"""The confusion-matrix kernel and the metric classes built on it, against the oracle (exact integer counts)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TASK = "AERIAL_LABEL-COSIA"


@pytest.mark.parametrize("C,shape", [(19, (16, 512, 512)), (19, (1, 7, 13)), (2, (3, 100)), (96, (2, 64, 64)), (5, (0,))])
def test_confusion_matrix_kernel_is_exact(cuda, C, shape):
    from oracle import metrics as om
    from flair_for_aigle_b200 import native as nv
    rng = np.random.default_rng(C + len(shape))
    n = int(np.prod(shape))
    # spatially coherent labels (runs of one class, like a real raster) with junk values mixed in
    runs = np.repeat(rng.integers(0, C, n // 50 + 1), 50)[:n]
    t = np.where(rng.random(n) < 0.03, rng.choice([-100, -1, C, 255]), runs).reshape(shape)
    p = np.where(rng.random(n) < 0.8, runs, rng.integers(-1, C + 1, n)).reshape(shape)
    tt, pp = torch.from_numpy(t).to(cuda), torch.from_numpy(p).to(cuda)
    cm = nv.confusion_matrix(tt, pp, C)
    ref = om.confusion_matrix(t, p, C)
    assert cm.dtype == torch.int64 and np.array_equal(cm.cpu().numpy(), ref)
    nv.confusion_matrix(tt, pp, C, out=cm)                       # accumulates
    assert np.array_equal(cm.cpu().numpy(), 2 * ref)
    with pytest.raises(nv.NativeError):
        nv.confusion_matrix(tt, pp, 97)


def test_jaccard_index_classes_vs_oracle(cuda):
    from oracle import metrics as om
    from flair_for_aigle_b200.flair_hub.tasks.metrics import MulticlassJaccardIndex
    rng = np.random.default_rng(5)
    C = 19
    metrics = {avg: MulticlassJaccardIndex(C, average=avg) for avg in ("weighted", None, "macro", "micro")}
    total = np.zeros((C, C), dtype=np.int64)
    for _ in range(3):
        t = rng.integers(0, 15, (2, 64, 64))                     # classes 15..18 never occur as labels
        p = np.where(rng.random(t.shape) < 0.6, t, rng.integers(0, C, t.shape))
        total += om.confusion_matrix(t, p, C)
        for m in metrics.values():
            m.update(torch.from_numpy(p).to(cuda).to(torch.int32), torch.from_numpy(t).to(cuda).to(torch.int32))
    for avg, m in metrics.items():
        got, ref = m.compute().cpu().numpy(), om.jaccard(total, avg)
        if avg is None:
            assert np.array_equal(got, ref)                      # exact counts, one float32 division per class
        else:
            assert np.isclose(got, ref, rtol=1e-6, atol=0), avg  # float32 reductions: summation order differs
    m = metrics["weighted"]
    m.reset()
    with pytest.raises(RuntimeError):
        m.compute()


def test_segmentation_task_metric_hooks(cuda, tmp_path):
    """validation_step / training_step feed the metrics like tasks_module.py:196-201,268-276; the epoch-end hooks return the
    values the reference logs, computed from the same confusion matrices the oracle counts on the host."""
    import bench
    from oracle import metrics as om
    from flair_for_aigle_b200.flair_hub.tasks.tasks_module import SegmentationTask
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, prepare_model_config
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster, register_raster
    from flair_for_aigle_b200.synthetic import synthetic_raster
    wpath = str(tmp_path / "w.safetensors")
    bench.make_weights(wpath, seed=7)
    name = "mem://metric_task"
    register_raster(name, ZoneRaster(synthetic_raster(512, 512, seed=1), 700000.0, 6600000.0, 0.2, name=name))
    cfg = inf.initialize_geometry_and_resolutions(bench.zonal_config(wpath, str(tmp_path), name, 2))
    cfg["device"] = cuda
    model = build_inference_model(cfg, {"AERIAL_RGBI": 512}).to(cuda)
    mcfg = prepare_model_config(cfg)
    mcfg["labels"] = [TASK]
    mcfg["labels_configs"] = {TASK: {"value_name": {k: f"c{k}" for k in range(19)}, "task_weight": 1.0,
                                     "value_weights": {"default": 1, "default_exceptions": {15: 0, 16: 0, 17: 0, 18: 0}}}}
    mcfg.setdefault("modalities", {}).setdefault("aux_loss", {})
    task = SegmentationTask(model, mcfg)
    g = torch.Generator(device="cpu").manual_seed(3)
    total = np.zeros((19, 19), dtype=np.int64)
    losses = []
    for _ in range(2):
        labels = torch.randint(0, 19, (2, 512, 512), generator=g)
        batch = {"AERIAL_RGBI": torch.randn(2, 4, 512, 512, generator=g).to(cuda),
                 TASK: torch.nn.functional.one_hot(labels, 19).permute(0, 3, 1, 2).float().to(cuda)}
        losses.append(float(task.validation_step(batch)))
        _, preds, targets = task.step(batch)
        assert torch.equal(targets[TASK].cpu(), labels.to(torch.int32))
        total += om.confusion_matrix(labels.numpy(), preds[TASK].cpu().numpy(), 19)
    out = task.on_validation_epoch_end()
    assert abs(out["val_loss"] - sum(losses) / 2) < 1e-6
    assert out["val_miou_COSIA"] == out["val_miou"] == pytest.approx(float(om.jaccard(total, "weighted")), rel=1e-6)
    per_class = om.jaccard(total, None)
    assert all(out[f"val_iou_COSIA_{k}_c{k}"] == float(per_class[k]) for k in range(19))
    assert task.val_metrics[TASK].confmat is None and task.val_loss.total is None          # reset

    task.configure_trainer({"optimizer": "adamw", "learning_rate": 1e-4, "optim_weight_decay": 0.01, "optim_betas": [0.9, 0.999]})
    small = {"AERIAL_RGBI": torch.randn(2, 4, 256, 256, generator=g).to(cuda),
             TASK: torch.randint(0, 19, (2, 256, 256), generator=g, dtype=torch.int32).to(cuda)}
    seen = np.zeros((19, 19), dtype=np.int64)
    tl = []
    for _ in range(2):
        loss, preds = task.training_step(small)
        tl.append(float(loss))
        seen += om.confusion_matrix(small[TASK].cpu().numpy(), preds[TASK].cpu().numpy(), 19)
    tr = task.on_train_epoch_end()
    assert tr["train_miou_COSIA"] == pytest.approx(float(om.jaccard(seen, "weighted")), rel=1e-6)
    assert abs(tr["train_loss"] - sum(tl) / 2) < 1e-6
    assert task.on_train_epoch_end() == {}

"""Training step, loss and optimizer side (tasks_module.py:133-167,377-391, module_setup.py:119-200): host logic
on CPU, the CUDA kernels against torch's CrossEntropyLoss / AdamW (what the reference calls) on the GPU."""
import numpy as np
import pytest
import torch

TASK = "AERIAL_LABEL-COSIA"
CFG5_WEIGHTS = {"default": 1, "default_exceptions": {15: 0, 16: 0, 17: 0, 18: 0}}     # BASELINE.json configs[4]


def _labels_config(n_cls=19, weights=CFG5_WEIGHTS, task_weight=1.0):
    return {"labels": [TASK], "labels_configs": {TASK: {"value_name": {i: f"c{i}" for i in range(n_cls)},
                                                       "value_weights": weights, "task_weight": task_weight}},
            "modalities": {"aux_loss": {}, "inputs": {"AERIAL_RGBI": True}}}


def test_flair_losses_weights_and_guards():
    from oracle.training import default_class_weights
    from flair_for_aigle_b200.flair_hub.tasks.module_setup import FLAIRLosses
    from flair_for_aigle_b200.flair_hub.tasks.tasks_module import SegmentationTask, init_optimizer
    cfg = _labels_config()
    losses = FLAIRLosses(cfg)
    w = losses.default_weights[TASK]
    assert torch.equal(w, default_class_weights(cfg["labels_configs"][TASK]))
    assert w.tolist() == [1.0] * 15 + [0.0] * 4 and set(losses.get_losses()) == {TASK}
    cfg_aux = _labels_config()
    cfg_aux["modalities"]["aux_loss"] = {"AERIAL_RGBI": True}
    with pytest.raises(NotImplementedError):
        FLAIRLosses(cfg_aux)
    with pytest.raises(ValueError, match="Unsupported optimizer type"):
        init_optimizer({"optimizer": "lion", "learning_rate": 1e-3}, [])
    with pytest.raises(NotImplementedError):
        init_optimizer({"optimizer": "sgd", "learning_rate": 1e-3}, [])
    with pytest.raises(NotImplementedError, match="backward"):
        SegmentationTask(model=None, config=cfg).step({}, training=True)


def test_oracle_step_is_the_weighted_mean():
    """The oracle's loss equals sum(w[t] nll) / sum(w[t]) written out by hand (what the CUDA kernel computes)."""
    from oracle.training import step
    torch.manual_seed(0)
    logits = torch.randn(2, 19, 8, 9)
    onehot = torch.nn.functional.one_hot(torch.randint(0, 19, (2, 8, 9)), 19).permute(0, 3, 1, 2).float()
    cfg = _labels_config(task_weight=0.5)
    loss, preds, targets = step(lambda b: ({TASK: logits}, {}), {TASK: onehot}, cfg)
    t = onehot.argmax(1)
    w = torch.tensor([1.0] * 15 + [0.0] * 4)[t]
    nll = torch.logsumexp(logits, 1) - logits.gather(1, t[:, None]).squeeze(1)
    assert torch.allclose(loss, 0.5 * (w * nll).sum() / w.sum(), rtol=1e-6)
    assert torch.equal(preds[TASK], logits.argmax(1)) and torch.equal(targets[TASK], t.int())


@pytest.mark.gpu
@pytest.mark.parametrize("shape,weights,task_weight", [((2, 19, 64, 96), CFG5_WEIGHTS, 1.0),
                                                       ((3, 13, 33, 47), {"default": 0.7, "default_exceptions": {2: 3.0}}, 0.25),
                                                       ((16, 19, 512, 512), CFG5_WEIGHTS, 1.0)])
def test_cross_entropy_forward_backward(cuda, shape, weights, task_weight):
    from oracle.training import default_class_weights
    from flair_for_aigle_b200 import native as nv
    from flair_for_aigle_b200.flair_hub.tasks.module_setup import WeightedCrossEntropy
    B, C, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + C)
    logits = (3 * torch.randn(shape, generator=g)).to(cuda)
    t = torch.randint(0, C, (B, H, W), generator=g).to(cuda)
    onehot = torch.nn.functional.one_hot(t, C).permute(0, 3, 1, 2).float().contiguous()
    targets = nv.onehot_argmax(onehot)
    assert torch.equal(targets.long(), t)
    w = default_class_weights({"value_name": list(range(C)), "value_weights": weights}).to(cuda)
    ref_logits = logits.clone().requires_grad_(True)
    ref = task_weight * torch.nn.CrossEntropyLoss(weight=w)(ref_logits, t)
    ref.backward()
    crit = WeightedCrossEntropy(w)
    loss, preds = crit(logits, targets, task_weight=task_weight, want_preds=True)
    assert abs(float(loss) - float(ref)) <= 2e-6 * abs(float(ref))
    assert torch.equal(preds.long(), torch.argmax(torch.softmax(logits, dim=1), dim=1))
    grad = crit.backward()
    scale = float(ref_logits.grad.abs().max())
    assert float((grad - ref_logits.grad).abs().max()) <= 2e-6 * scale
    again, _ = crit(logits, targets, task_weight=task_weight, want_preds=True)
    assert float(again) == float(loss)                                        # deterministic reduction


@pytest.mark.gpu
def test_cross_entropy_ignores_out_of_range_targets(cuda):
    """Round-1 advisor finding: integer targets of -100 (nn.CrossEntropyLoss's ignore_index), negative or >= C used to index
    the class weights out of bounds.  They are ignored now: no contribution to the loss or to sum(w), zero gradient --
    exactly what torch does for -100."""
    from flair_for_aigle_b200.flair_hub.tasks.module_setup import WeightedCrossEntropy
    B, C, H, W = 2, 19, 40, 56
    g = torch.Generator(device="cpu").manual_seed(5)
    logits = (3 * torch.randn((B, C, H, W), generator=g)).to(cuda)
    t = torch.randint(0, C, (B, H, W), generator=g)
    holes = torch.rand((B, H, W), generator=g) < 0.3
    t_torch = t.clone()
    t_torch[holes] = -100
    t_ours = t.clone()
    t_ours[holes] = torch.tensor([-100, -1, C, 1000])[torch.randint(0, 4, (int(holes.sum()),), generator=g)]
    w = (torch.rand(C, generator=g) + 0.5).to(cuda)
    ref_logits = logits.clone().requires_grad_(True)
    ref = torch.nn.CrossEntropyLoss(weight=w)(ref_logits, t_torch.to(cuda))
    ref.backward()
    crit = WeightedCrossEntropy(w)
    loss = crit(logits, t_ours.to(cuda).int())
    grad = crit.backward()
    assert abs(float(loss) - float(ref)) <= 2e-6 * abs(float(ref))
    assert float((grad - ref_logits.grad).abs().max()) <= 2e-6 * float(ref_logits.grad.abs().max())
    assert float(grad.permute(0, 2, 3, 1)[holes.to(cuda)].abs().max()) == 0.0


@pytest.mark.gpu
def test_adamw_matches_torch(cuda):
    from oracle.training import init_optimizer as oracle_opt
    from flair_for_aigle_b200.flair_hub.tasks.tasks_module import init_optimizer
    cfg = {"optimizer": "adamw", "learning_rate": 5e-5, "optim_weight_decay": 0.01, "optim_betas": [0.9, 0.999]}   # configs[4]
    g = torch.Generator(device="cpu").manual_seed(3)
    shapes = [(128, 4, 4, 4), (128,), (512, 128), (19, 16, 3, 3), (1,), (1000003,)]
    init = [torch.randn(s, generator=g).to(cuda) for s in shapes]
    ref_params = [p.clone().requires_grad_(True) for p in init]
    ours = [p.clone() for p in init]
    ref = oracle_opt(cfg, ref_params)
    opt = init_optimizer(cfg, ours)
    for step in range(6):
        grads = [(torch.randn(s, generator=g) * (10.0 if step == 2 else 1.0)).to(cuda) for s in shapes]
        for p, gr, og in zip(ref_params, grads, opt.grads):
            p.grad = gr.clone()
            og.copy_(gr)
        ref.step()
        opt.step()
        for p, q in zip(ref_params, ours):
            assert q.data_ptr() >= opt.arena.data_ptr()                       # parameters alias the arena
            d = (p.detach() - q).abs().max().item()
            assert d <= 1e-6 * max(1.0, p.detach().abs().max().item()), (step, d)
    st = ref.state[ref_params[2]]
    off = sum(int(np.prod(s)) for s in shapes[:2])
    n = int(np.prod(shapes[2]))
    assert torch.allclose(opt.exp_avg[off:off + n].view(shapes[2]), st["exp_avg"], rtol=1e-5, atol=1e-9)
    assert torch.allclose(opt.exp_avg_sq[off:off + n].view(shapes[2]), st["exp_avg_sq"], rtol=1e-5, atol=1e-12)


@pytest.mark.gpu
def test_validation_step_against_the_oracle(cuda, tmp_path):
    """SegmentationTask.step(batch, training=False) (tasks_module.py:133-167) through the zonal engine vs the oracle model +
    torch CrossEntropyLoss on the same seeded weights: the loss within 1 %, predictions >= 98.5 % equal."""
    import bench
    from safetensors.torch import load_file
    from oracle.models import FlairHubOracle
    from oracle.training import step as oracle_step
    from flair_for_aigle_b200.flair_hub.tasks.tasks_module import SegmentationTask
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster, register_raster
    from flair_for_aigle_b200.synthetic import synthetic_raster
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    wpath = str(tmp_path / "w.safetensors")
    bench.make_weights(wpath, seed=7)
    name = "mem://train_step"
    register_raster(name, ZoneRaster(synthetic_raster(512, 512, seed=1), 700000.0, 6600000.0, 0.2, name=name))
    cfg = bench.zonal_config(wpath, str(tmp_path), name, 4)
    cfg = inf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = cuda
    model = build_inference_model(cfg, compute_patch_sizes(cfg)).to(cuda)
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(2, 4, 512, 512, generator=g).to(cuda)
    onehot = torch.nn.functional.one_hot(torch.randint(0, 19, (2, 512, 512), generator=g), 19).permute(0, 3, 1, 2).float().to(cuda)
    lc = _labels_config()
    task = SegmentationTask(model, lc)
    loss, preds, targets = task.step({"AERIAL_RGBI": x, TASK: onehot}, training=False)
    oracle = FlairHubOracle("convnextv2_base-unet", {"AERIAL_RGBI": 4}, {TASK: 19}).eval()
    oracle.load_state_dict(load_file(wpath), strict=True)
    oracle = oracle.to(cuda)
    with torch.no_grad():
        ref_loss, ref_preds, ref_targets = oracle_step(oracle, {"AERIAL_RGBI": x, TASK: onehot}, lc)
    assert torch.equal(targets[TASK], ref_targets[TASK])
    assert abs(float(loss) - float(ref_loss)) <= 1e-2 * abs(float(ref_loss))
    agree = (preds[TASK] == ref_preds[TASK]).float().mean().item()
    print(f"validation step: loss {float(loss):.5f} vs oracle {float(ref_loss):.5f}, predictions agree {agree:.4f}")
    assert agree >= 0.985
    grads = task.loss_gradients()[TASK]
    assert tuple(grads.shape) == (2, 19, 512, 512) and bool(torch.isfinite(grads).all())
    assert abs(float(grads.sum())) < 1e-3                                     # softmax - onehot sums to zero per pixel


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(4096, 512, 128), (16384, 2048, 512), (1024, 128, 512), (640, 64, 64)])
def test_linear_backward(cuda, M, N, K):
    """dX, dW, db of a Linear layer (the ConvNeXt MLPs, 1x1 convolutions, fusion conv_f) on the tcgen05 GEMM through explicit
    operand transposes, vs fp32 matmuls of the same bf16 operands: dX within bf16 rounding, dW / db fp32-accumulated."""
    from flair_for_aigle_b200 import native as nv
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    X = torch.randn(M, K, generator=g).bfloat16().to(cuda)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(cuda)
    dY = torch.randn(M, N, generator=g).bfloat16().to(cuda)
    assert torch.equal(nv.transpose_bf16(X), X.t().contiguous())
    dX, dW, db = nv.linear_backward(dY, X, W)
    rX = dY.float() @ W.float()
    rW = dY.float().t() @ X.float()
    rb = dY.float().sum(0)
    assert float((dX.float() - rX).abs().max()) <= 1e-2 * float(rX.abs().max())
    assert float((dW - rW).abs().max()) <= 2e-3 * float(rW.abs().max())
    assert float((db - rb).abs().max()) <= 1e-4 * max(1.0, float(rb.abs().max()))
    dX2, dW2, db2 = nv.linear_backward(dY, X, W)
    assert torch.equal(dW, dW2) and torch.equal(db, db2) and torch.equal(dX, dX2)      # deterministic


def test_module_factories_like_the_reference():
    """module_setup.py:47-117: ``build_segmentation_module`` (train: with the losses; predict: without, from a config that
    carries no class weights) and ``get_input_img_sizes`` (first batch of the data module's loader)."""
    import bench
    from flair_for_aigle_b200.flair_hub.tasks.module_setup import build_segmentation_module, get_input_img_sizes
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    zc = bench.zonal_config("w", "/tmp", "unused", 1)
    zc["monotemp_arch"] = "resnet34-unet"
    cfg = prepare_model_config(zc)
    task = build_segmentation_module(cfg, {"AERIAL_RGBI": 512}, "predict")
    assert type(task).__name__ == "SegmentationTask" and type(task.model).__name__ == "FLAIR_HUB_Model" and task.criterion is None
    label = cfg["labels"][0]
    cfg["labels_configs"][label]["value_weights"] = {"default": 1, "default_exceptions": {18: 0}}
    task = build_segmentation_module(cfg, {"AERIAL_RGBI": 512}, "train")
    assert list(task.criterion) == [label] and float(task.criterion[label].weight[18]) == 0.0
    with pytest.raises(AssertionError, match="stage"):
        build_segmentation_module(cfg, {"AERIAL_RGBI": 512}, "fit")

    class DM:
        def setup(self, stage):
            self.stage = stage

        def train_dataloader(self):
            return [{"AERIAL_RGBI": torch.zeros(2, 4, 512, 512), "DEM_ELEV": torch.zeros(2, 1, 128, 128), "other": 1}]

        def predict_dataloader(self):
            return [{"AERIAL_RGBI": torch.zeros(2, 4, 256, 256)}]
    c = {"modalities": {"inputs": {"AERIAL_RGBI": True, "DEM_ELEV": True, "SPOT_RGBI": False}}}
    assert get_input_img_sizes(c, DM(), "fit") == {"AERIAL_RGBI": 512, "DEM_ELEV": 128}
    assert get_input_img_sizes(c, DM(), "predict") == {"AERIAL_RGBI": 256}

"""Where does the training step's gradient error come from?  (test infrastructure: uses the oracle)

The engine's gradients differ from fp32 autograd (whole-model cosine 0.985-0.990, tests/test_gpu_train_step.py).  Two
sources: the FORWARD activations and weights are rounded to bf16 before every tensor-core product (so the backward starts
from slightly different ReLU masks, GELU slopes, normalisation statistics), and the BACKWARD signals are rounded to bf16 too.
This script isolates the first: it runs the fp32 oracle in training mode with the engine's forward rounding points switched
on through straight-through hooks (value rounded, gradient passed unchanged), for bf16 and for fp16, and reports the
whole-model gradient cosine against the unrounded run.  If the forward rounding alone explains the engine's figure, an
fp16 forward (the inference engine's format) would lift it; if not, the backward's own rounding is the floor.

    python tests/diag/grad_precision_budget.py [--tile 128] [--batch 2] [--device cpu|cuda]"""
import argparse
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.models import (ConvNeXtBlock, FlairHubOracle, GlobalResponseNorm, LayerNorm2d, randomize_)  # noqa: E402
from oracle.training import step as oracle_step  # noqa: E402

TASK = "AERIAL_LABEL-COSIA"


def ste_round(dt):
    def f(t):
        return t + (t.to(dt).float() - t).detach()
    return f


def install(model: nn.Module, dt, weights: bool, acts: bool):
    """Rounding points of engine/convnext_train.py: LayerNorm outputs (a1, downsample inputs), fc1 output (h), GELU output
    (g), GRN output (a2), every decoder Conv-BN-ReLU output and the stage outputs handed to the decoder; GEMM / convolution
    weights.  Returns the hook handles and the original weights."""
    q = ste_round(dt)
    handles, saved = [], {}
    if acts:
        for m in model.modules():
            if isinstance(m, (nn.LayerNorm, LayerNorm2d, nn.GELU, GlobalResponseNorm, nn.ReLU)):
                handles.append(m.register_forward_hook(lambda mod, inp, out: q(out)))
            elif isinstance(m, nn.Linear) and m.out_features > m.in_features:           # fc1: the pre-GELU tensor h
                handles.append(m.register_forward_hook(lambda mod, inp, out: q(out)))
            elif isinstance(m, ConvNeXtBlock):                                           # nothing: the residual stream is fp32
                pass
    if weights:
        for name, p in model.named_parameters():
            if p.dim() >= 2 and "conv_dw" not in name:                                   # depthwise taps stay fp32 in the engine
                saved[name] = p.data.clone()
                p.data = p.data.to(dt).float()
    return handles, saved


def grads(model, batch, cfg):
    model.zero_grad(set_to_none=True)
    loss, _, _ = oracle_step(model, batch, cfg)
    loss.backward()
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}, float(loss)


def cosine(a, b):
    dot = sum(float((a[k].double() * b[k].double()).sum()) for k in a)
    na = sum(float((a[k].double() ** 2).sum()) for k in a) ** 0.5
    nb = sum(float((b[k].double() ** 2).sum()) for k in a) ** 0.5
    worst = min((float(torch.nn.functional.cosine_similarity(a[k].flatten().double(), b[k].flatten().double(), dim=0)), k) for k in a)
    return dot / (na * nb), worst


CASES = (("bf16 activations + weights", torch.bfloat16, True, True), ("bf16 activations only", torch.bfloat16, False, True),
         ("bf16 weights only", torch.bfloat16, True, False), ("fp16 activations + weights", torch.float16, True, True))


def budget(arch: str = "convnextv2_base-unet", tile: int = 128, batch_size: int = 2, device: str = "cpu", cases=CASES):
    """-> (loss of the unrounded run, [(label, loss, whole-model cosine, (worst cosine, tensor name))])."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device(device)
    mods = {"AERIAL_RGBI": 4, "DEM_ELEV": 1}
    model = FlairHubOracle(arch, mods, {TASK: 19})
    randomize_(model, seed=3)
    model = model.to(dev).train()
    cfg = {"labels_configs": {TASK: {"value_name": list(range(19)), "task_weight": 1.0,
                                     "value_weights": {"default": 1, "default_exceptions": {15: 0, 16: 0, 17: 0, 18: 0}}}}}
    g = torch.Generator(device="cpu").manual_seed(17)
    batch = {m: torch.randn(batch_size, c, tile, tile, generator=g).to(dev) for m, c in mods.items()}
    batch[TASK] = torch.nn.functional.one_hot(torch.randint(0, 19, (batch_size, tile, tile), generator=g), 19).permute(0, 3, 1, 2).float().to(dev)
    ref, loss0 = grads(model, batch, cfg)
    rows = []
    for label, dt, w, ac in cases:
        handles, saved = install(model, dt, w, ac)
        got, loss = grads(model, batch, cfg)
        for h in handles:
            h.remove()
        for name, p in model.named_parameters():
            if name in saved:
                p.data = saved[name]
        c, worst = cosine(got, ref)
        rows.append((label, loss, c, worst))
    return loss0, rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--arch", default="convnextv2_base-unet")
    a = ap.parse_args()
    loss0, rows = budget(a.arch, a.tile, a.batch, a.device)
    print(f"forward-only rounding (straight-through), {a.arch} 2 encoders, batch {a.batch} x {a.tile}^2, loss {loss0:.5f}")
    print(f"{'rounded in the forward':44s} {'loss':>9s} {'whole-model cosine':>19s}   worst tensor")
    for label, loss, c, worst in rows:
        print(f"{label:44s} {loss:9.5f} {c:19.5f}   {worst[0]:.4f} {worst[1]}")


if __name__ == "__main__":
    main()

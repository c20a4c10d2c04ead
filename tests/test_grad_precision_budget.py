"""The experiment behind the training step's fp16 forward, at a size the CPU suite can afford: with the
engine's forward rounding points switched on in the fp32 oracle (values rounded, gradients passed straight through), bf16
rounding of the forward tensors alone moves the whole-model gradient measurably away from the unrounded one, fp16 rounding
several times less."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "diag"))


def test_forward_rounding_explains_the_gradient_error():
    import grad_precision_budget as gb
    cases = (("bf16", torch.bfloat16, True, True), ("fp16", torch.float16, True, True))
    loss0, rows = gb.budget("convnextv2_atto-unet", tile=64, batch_size=2, device="cpu", cases=cases)
    got = {label: (loss, cos, worst) for label, loss, cos, worst in rows}
    print(f"loss {loss0:.5f}; bf16 forward: cosine {got['bf16'][1]:.5f} worst {got['bf16'][2][0]:.4f}; "
          f"fp16 forward: cosine {got['fp16'][1]:.5f} worst {got['fp16'][2][0]:.4f}")
    err = lambda c: (1.0 - c * c) ** 0.5                       # relative size of the gradient error
    assert got["fp16"][1] > got["bf16"][1]
    assert err(got["fp16"][1]) < 0.5 * err(got["bf16"][1])     # three more significand bits: measured ~ 3 x smaller error
    assert got["fp16"][1] > 0.995
    assert abs(got["fp16"][0] - loss0) < 1e-4 * abs(loss0)

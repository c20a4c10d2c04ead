"""Backward of one ConvNeXt-V2 block (first slice of the model backward) against torch autograd on the oracle's
ConvNeXtBlock with the same bf16-exact weights.  The product path keeps the inference engine's number formats (bf16 GEMM
operands, fp32 accumulation, fp32 residual stream), so gradients are compared by direction and relative error."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    err = ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()
    return cos, err


@pytest.mark.parametrize("B,H,C", [(2, 32, 128), (3, 16, 256), (2, 16, 512)])
def test_convnext_block_forward_backward_vs_autograd(cuda, B, H, C):
    from oracle.models import ConvNeXtBlock
    from flair_for_aigle_b200.engine.convnext_train import ConvNeXtBlockTrain
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(C + H)
    blk = ConvNeXtBlock(C)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith("grn.weight") or n.endswith("grn.bias"):
                p.copy_(torch.randn_like(p) * 0.5)
            elif n == "norm.weight":
                p.copy_(1.0 + 0.2 * torch.randn_like(p))
            elif p.dim() == 1:
                p.copy_(0.1 * torch.randn_like(p))
            elif n == "conv_dw.weight":
                p.copy_(torch.randn_like(p) / 7.0)
            else:
                p.copy_((torch.randn_like(p) / p.shape[1] ** 0.5))
            p.copy_(p.bfloat16().float())                       # bf16-exact weights: both sides see the same numbers
    blk = blk.to(cuda)
    x = torch.randn(B, H, H, C, device=cuda)
    dy = torch.randn(B, H, H, C, device=cuda)

    xr = x.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    yr = blk(xr)
    yr.backward(dy.permute(0, 3, 1, 2).contiguous())
    ref_y = yr.detach().permute(0, 2, 3, 1)
    ref_dx = xr.grad.permute(0, 2, 3, 1)

    eng = ConvNeXtBlockTrain({n: p.detach() for n, p in blk.named_parameters()})
    y = eng.forward(x)
    dx, grads = eng.backward(dy)
    torch.cuda.synchronize()
    cos, err = _rel(y, ref_y)
    print(f"forward: cos {cos:.6f} max rel err {err:.4f}")
    assert cos > 0.9999 and err < 2e-2
    cos, err = _rel(dx, ref_dx)
    print(f"dx: cos {cos:.6f} max rel err {err:.4f}")
    assert cos > 0.9995 and err < 3e-2
    for n, p in blk.named_parameters():
        cos, err = _rel(grads[n], p.grad)
        print(f"{n:18s} cos {cos:.6f} max rel err {err:.4f}")
        assert tuple(grads[n].shape) == tuple(p.grad.shape)
        assert cos > 0.999 and err < 5e-2, n
    # deterministic: same inputs, same bits
    y2 = eng.forward(x)
    dx2, grads2 = eng.backward(dy)
    assert torch.equal(y, y2) and torch.equal(dx, dx2)
    assert all(torch.equal(grads[k], grads2[k]) for k in grads)

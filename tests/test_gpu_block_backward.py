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


@pytest.mark.parametrize("depths,B,P,cin", [((1, 1, 2, 1), 2, 128, 4), ((3, 3, 27, 3), 1, 128, 1)])
def test_convnext_encoder_backward_vs_autograd(cuda, depths, B, P, cin):
    """The whole ConvNeXt-V2 feature extractor (stem, 4 stages, downsample layers; base widths, the second case with the real
    base depths and the 1-channel elevation stem): gradients of every parameter for a loss that touches all four outputs."""
    from oracle.models import ConvNeXtV2Features
    from flair_for_aigle_b200.engine.convnext_train import ConvNeXtV2EncoderTrain
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dims = (128, 256, 512, 1024)
    torch.manual_seed(sum(depths) + cin)
    enc = ConvNeXtV2Features(cin, depths, dims)
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if "grn" in n:
                p.copy_(torch.randn_like(p) * 0.3)
            elif n.endswith("norm.weight") or n.endswith("stem_1.weight") or n.endswith("downsample.0.weight"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn_like(p))
            else:
                fan_in = p[0].numel()
                p.copy_(torch.randn_like(p) / fan_in ** 0.5)
            p.copy_(p.bfloat16().float())
    enc = enc.to(cuda)
    x = torch.randn(B, cin, P, P, device=cuda)
    feats_ref = enc(x)
    dfe = [torch.randn_like(f) / f[0].numel() ** 0.5 for f in feats_ref]           # NCHW
    torch.autograd.backward(feats_ref, dfe)

    eng = ConvNeXtV2EncoderTrain({n: p.detach() for n, p in enc.named_parameters()}, depths, dims)
    feats = eng.forward(x)
    for f, r in zip(feats, feats_ref):
        cos, err = _rel(f, r.detach().permute(0, 2, 3, 1))
        assert cos > 0.999 and err < 6e-2, (cos, err)
    grads = eng.backward([d.permute(0, 2, 3, 1).contiguous() for d in dfe])
    torch.cuda.synchronize()
    worst = (1.0, "", 0.0)
    names = [n for n, _ in enc.named_parameters()]
    assert sorted(grads) == sorted(names)
    for n, p in enc.named_parameters():
        cos, err = _rel(grads[n], p.grad)
        if cos < worst[0]:
            worst = (cos, n, err)
        assert tuple(grads[n].shape) == tuple(p.grad.shape), n
        assert cos > 0.99, (n, cos, err)
    print(f"encoder depths {depths}: {len(names)} parameter gradients, worst cosine {worst[0]:.5f} at {worst[1]} (max rel err {worst[2]:.3f})")

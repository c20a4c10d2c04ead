"""GPU parity: tcgen05 GEMM (+ fused epilogues) vs an fp32 torch reference and the SIMT kernel, for both 16-bit operand
formats the kernel takes per call (float16 = the inference engines, bfloat16 = the training step)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DTYPES = [torch.float16, torch.bfloat16]
# 16-bit output rounding: 2^-11 relative for float16, 2^-8 for bfloat16 (x safety factor), else fp32 accumulation-order noise
OUT_TOL = {torch.float16: 3e-3, torch.bfloat16: 2e-2, torch.float32: 2e-4}


def _ref(A, B, bias, mode, resid, rows_per_sample):
    from flair_for_aigle_b200 import native as nv
    Af, Bf = A.float(), B.float()
    if B.dim() == 3:
        nb = B.shape[0]
        acc = torch.cat([Af[i * rows_per_sample:(i + 1) * rows_per_sample] @ Bf[i].t() for i in range(nb)])
    else:
        acc = Af @ Bf.t()
    v = acc + bias
    sumsq = None
    if mode == nv.EPI_GELU_SUMSQ:
        v = torch.nn.functional.gelu(v)
        # GRN statistics are taken from the bf16 values fc2 will consume (the kernel re-reads its staged output tile)
        sumsq = (v.to(A.dtype).float().view(A.shape[0] // 128, 128, -1) ** 2).sum(1)
    elif mode == nv.EPI_GELU_BF16:
        v = torch.nn.functional.gelu(v)
    elif mode == nv.EPI_RELU_BF16:
        v = torch.relu(v)
    elif mode == nv.EPI_RESID_F32:
        v = v + resid
    return v, sumsq


@pytest.mark.parametrize("impl", ["tcgen05", "simt"])
@pytest.mark.parametrize("M,N,K,rps,bb", [
    (256, 128, 64, 128, 1), (512, 256, 128, 256, 1), (1024, 512, 512, 256, 4),
    (2048, 2048, 512, 1024, 1), (384, 64, 192, 128, 1), (2048, 512, 2048, 1024, 2),
])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_modes(cuda, impl, M, N, K, rps, bb, mode, dt):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(M + N + K + mode)
    torch.backends.cuda.matmul.allow_tf32 = False
    A = (torch.randn(M, K, device=cuda) * 0.5).to(dt)
    B = (torch.randn((bb, N, K) if bb > 1 else (N, K), device=cuda) / K ** 0.5).to(dt)
    bias = torch.randn(N, device=cuda) * 0.1
    resid = torch.randn(M, N, device=cuda) if mode == nv.EPI_RESID_F32 else None
    sumsq = torch.full((M // 128, N), -1.0, device=cuda) if mode == nv.EPI_GELU_SUMSQ else None
    out = nv.gemm_bf16(A, B, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps, impl=impl)
    torch.cuda.synchronize()
    ref, ref_sq = _ref(A, B, bias, mode, resid, rps)
    err = (out.float() - ref).abs().max().item()
    assert out.dtype == (torch.float32 if mode in (nv.EPI_RESID_F32, nv.EPI_F32) else dt)
    tol = OUT_TOL[out.dtype]
    assert err < tol * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    if sumsq is not None:
        rel = ((sumsq - ref_sq).abs() / ref_sq.clamp_min(1e-3)).max().item()
        assert rel < 5e-3, f"sumsq rel err {rel}"   # tanh.approx (2^-11 rel.) moves ~10 % of the bf16 roundings


@pytest.mark.parametrize("M,N,K,rps,bb", [
    (512, 256, 128, 256, 1), (2048, 2048, 512, 1024, 1), (2048, 512, 2048, 1024, 2),
    (1408, 512, 256, 1408, 1),          # last pair tile: only the leader's 128 rows exist
    (10240, 1024, 256, 1024, 10),       # 160 pair tiles over 74 pairs: ring + TMEM double buffering across tiles
])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_pair_kernel(cuda, monkeypatch, M, N, K, rps, bb, mode, dt):
    """cta_group::2 kernel (gemm_tcgen05_2sm.cu), forced for every N % 256 == 0 shape."""
    from flair_for_aigle_b200 import native as nv
    monkeypatch.setenv("FZ_GEMM_PAIR", "2")
    torch.manual_seed(M + N + K + mode)
    torch.backends.cuda.matmul.allow_tf32 = False
    A = (torch.randn(M, K, device=cuda) * 0.5).to(dt)
    B = (torch.randn((bb, N, K) if bb > 1 else (N, K), device=cuda) / K ** 0.5).to(dt)
    bias = torch.randn(N, device=cuda) * 0.1
    resid = torch.randn(M, N, device=cuda) if mode == nv.EPI_RESID_F32 else None
    sumsq = torch.full((M // 128, N), -1.0, device=cuda) if mode == nv.EPI_GELU_SUMSQ else None
    out = nv.gemm_bf16(A, B, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps)
    torch.cuda.synchronize()
    ref, ref_sq = _ref(A, B, bias, mode, resid, rps)
    err = (out.float() - ref).abs().max().item()
    tol = OUT_TOL[out.dtype]
    assert err < tol * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    if sumsq is not None:
        rel = ((sumsq - ref_sq).abs() / ref_sq.clamp_min(1e-3)).max().item()
        assert rel < 5e-3, f"sumsq rel err {rel}"   # tanh.approx (2^-11 rel.) moves ~10 % of the bf16 roundings
    # same inputs, same launch -> identical bits (no atomics anywhere)
    out2 = nv.gemm_bf16(A, B, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_inplace_residual(cuda, dt):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    M, N, K = 1024, 256, 1024
    A = torch.randn(M, K, device=cuda).to(dt)
    B = (torch.randn(N, K, device=cuda) / 32).to(dt)
    bias = torch.randn(N, device=cuda)
    x = torch.randn(M, N, device=cuda)
    ref = A.float() @ B.float().t() + bias + x
    nv.gemm_bf16(A, B, nv.EPI_RESID_F32, bias=bias, resid=x, out=x)
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("dt", DTYPES)
def test_gelu_fast_matches_erf(cuda, dt):
    """The epilogue's SFU GELU vs torch's erf GELU: K=64 identity-ish GEMM isolates it."""
    from flair_for_aigle_b200 import native as nv
    M, N, K = 4096, 64, 64
    x = torch.linspace(-12, 12, M * N, device=cuda).view(M, N)
    A = torch.zeros(M, K, device=cuda)
    A[:, 0] = 1.0
    B = torch.zeros(N, K, device=cuda)
    # acc = 0 -> value comes from the bias path only per column; use bias sweep instead
    bias = torch.linspace(-10, 10, N, device=cuda)
    sumsq = torch.zeros(M // 128, N, device=cuda)
    out = nv.gemm_bf16(A.to(dt), B.to(dt), nv.EPI_GELU_SUMSQ, bias=bias, sumsq=sumsq, rows_per_sample=128)
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(bias)
    # fitted GELU (2.6e-5 abs) + tanh.approx (2^-11 rel.) + one output rounding
    bound = (2 ** -8 if dt == torch.bfloat16 else 2 ** -9) * ref.abs().max().item()
    assert (out.float() - ref[None, :]).abs().max().item() <= bound


def test_fp16_output_saturates_instead_of_overflowing(cuda):
    """|value| > 65504 is stored as +-65504 (cvt.rn.satfinite), never inf."""
    from flair_for_aigle_b200 import native as nv
    M, N, K = 256, 64, 64
    A = torch.zeros(M, K, device=cuda, dtype=torch.float16)
    B = torch.zeros(N, K, device=cuda, dtype=torch.float16)
    bias = torch.linspace(-2.0e5, 2.0e5, N, device=cuda)
    out = nv.gemm_bf16(A, B, nv.EPI_BF16, bias=bias)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert torch.equal(out[0].float(), bias.clamp(-65504.0, 65504.0).half().float())


def test_gemm_rejects_mixed_operand_formats(cuda):
    from flair_for_aigle_b200 import native as nv
    A = torch.zeros(256, 64, device=cuda, dtype=torch.float16)
    B = torch.zeros(64, 64, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(nv.NativeError):
        nv.gemm_bf16(A, B, nv.EPI_BF16)
    with pytest.raises(nv.NativeError):
        nv.gemm_bf16(A, B.half(), nv.EPI_F32, out=torch.zeros(256, 64, device=cuda, dtype=torch.float16))


@pytest.mark.parametrize("pair", ["0", "2"])
@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_reverse_tile_order_is_bit_identical(cuda, monkeypatch, pair, dt):
    """FZ_EPI_REVERSE_TILES only changes which CTA computes which tile."""
    from flair_for_aigle_b200 import native as nv
    monkeypatch.setenv("FZ_GEMM_PAIR", pair)
    torch.manual_seed(3)
    M, N, K = 4096 + 128, 512, 512
    A = torch.randn(M, K, device=cuda).to(dt)
    B = (torch.randn(N, K, device=cuda) / 20).to(dt)
    bias = torch.randn(N, device=cuda)
    for mode in (nv.EPI_BF16, nv.EPI_GELU_BF16, nv.EPI_F32):
        a = nv.gemm_bf16(A, B, mode, bias=bias)
        b = nv.gemm_bf16(A, B, mode | nv.EPI_REVERSE_TILES, bias=bias)
        torch.cuda.synchronize()
        assert torch.equal(a, b)
    sq1 = torch.zeros(M // 128, N, device=cuda)
    sq2 = torch.zeros(M // 128, N, device=cuda)
    nv.gemm_bf16(A, B, nv.EPI_GELU_SUMSQ, bias=bias, sumsq=sq1)
    nv.gemm_bf16(A, B, nv.EPI_GELU_SUMSQ | nv.EPI_REVERSE_TILES, bias=bias, sumsq=sq2)
    torch.cuda.synchronize()
    assert torch.equal(sq1, sq2)


@pytest.mark.parametrize("M,N,K,splits", [(512, 128, 16384, None), (64, 320, 65536, None), (2048, 512, 16384, None),
                                         (128, 64, 4096, 7), (256, 128, 1024, 1), (1024, 256, 8192, 64), (16, 64, 262144, None)])
@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_splitk(cuda, M, N, K, splits, dt):
    """Small output, long reduction (weight gradients): split-K pieces inside one launch, added in a fixed order."""
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(M + N)
    torch.backends.cuda.matmul.allow_tf32 = False
    A = (torch.randn(M, K, device=cuda) * 0.5).to(dt)
    B = (torch.randn(N, K, device=cuda) * 0.5).to(dt)
    out = nv.gemm_splitk(A, B, splits=splits)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    sd = ref.std().item()
    err = (out.double() - ref).abs().max().item()
    # fp32 accumulation in TMEM over K / 16 MMA steps (the tensor core truncates, so the error of one long chain grows
    # ~linearly: 4e-4 of the output spread at K = 262144; split-K shortens the chains)
    assert err < 5e-4 * sd, (err, sd)
    assert torch.equal(out, nv.gemm_splitk(A, B, splits=splits))              # deterministic
    used = splits if splits is not None else nv.lib().fz_gemm_splitk_max_splits(M, N, K)
    assert 1 <= used <= 64
    if splits is None and K >= 16384:
        assert used > 1
    plain = nv.gemm_bf16(A, B, nv.EPI_F32)
    err_plain = (plain.double() - ref).abs().max().item()
    print(f"M={M} N={N} K={K} {dt}: split-K x{used} max err {err / sd:.2e} of the output spread, one chain {err_plain / sd:.2e}")
    assert err_plain < 2e-3 * sd


@pytest.mark.parametrize("M,N,K,splits", [(256, 128, 64, 1), (512, 128, 16384, None), (64, 320, 65536, None), (2048, 512, 16384, None),
                                         (128, 64, 4096, 7), (16, 64, 262144, None), (72, 192, 640, 1), (1024, 256, 8192, 64)])
@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_splitk_transposed_operands_in_place(cuda, M, N, K, splits, dt):
    """out = At^T Bt with At [K,M], Bt [K,N] read in place (MN-major tcgen05 descriptors): the weight-gradient GEMM dW = dY^T X
    without transposed copies.  Same values as the K-major split-K kernel on explicitly transposed operands."""
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(M + N + 1)
    torch.backends.cuda.matmul.allow_tf32 = False
    At = (torch.randn(K, M, device=cuda) * 0.5).to(dt)
    Bt = (torch.randn(K, N, device=cuda) * 0.5).to(dt)
    out = nv.gemm_splitk_tn(At, Bt, splits=splits)
    torch.cuda.synchronize()
    ref = At.double().t() @ Bt.double()
    sd = ref.std().item()
    assert (out.double() - ref).abs().max().item() < 5e-4 * sd
    assert torch.equal(out, nv.gemm_splitk_tn(At, Bt, splits=splits))                      # deterministic
    if M % 8 == 0 and K % 64 == 0:
        used = splits if splits is not None else nv.lib().fz_gemm_splitk_max_splits(M, N, K)
        same = nv.gemm_splitk(At.t().contiguous(), Bt.t().contiguous(), splits=used)      # K-major operands, same pieces
        assert torch.equal(out, same)
    # the helper the training step calls (pads the rows to a multiple of 64)
    dW = nv.weight_gradient(At[:K - 3], Bt[:K - 3])
    ref2 = At[:K - 3].double().t() @ Bt[:K - 3].double()
    assert (dW.double() - ref2).abs().max().item() < 5e-4 * max(ref2.std().item(), 1e-6)


@pytest.mark.parametrize("M,N,K", [(512, 128, 16384), (64, 320, 65536), (2048, 512, 4096), (16, 64, 8192)])
def test_weight_gradient_bf16_gradients_times_fp16_activations(cuda, M, N, K):
    """The training weight gradient dW = dY^T X with dY in bf16 (gradient range) and X in fp16 (the forward activations'
    format).  One tcgen05 MMA takes one operand format (a descriptor with a_format = BF16 and b_format = F16 was tried: illegal
    instruction), so ``weight_gradient`` re-rounds X to bf16 (fz_cast_f16_bf16) and the GEMM itself rejects the mix."""
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(M + N)
    dY = (torch.randn(K, M, device=cuda) * 0.5).to(torch.bfloat16)
    X = (torch.randn(K - 3, N, device=cuda) * 0.5).to(torch.float16)
    dW = nv.weight_gradient(dY[:K - 3], X)
    ref = dY[:K - 3].double().t() @ X.to(torch.bfloat16).double()            # X rounded to bf16: what the kernel multiplies
    assert (dW.double() - ref).abs().max().item() < 5e-4 * ref.std().item()
    exact = dY[:K - 3].double().t() @ X.double()
    print(f"[{M}x{N}] K={K}: re-rounding the activations to bf16 moves dW by {(ref - exact).abs().max().item() / exact.std().item():.1e} of its std")
    with pytest.raises(nv.NativeError):
        nv.gemm_splitk_tn(dY[:K - 3], X)

"""First-light diagnostics + micro-benchmark for the tcgen05 GEMM (run on the B200 box).
Writes gpurun_out/gemm_check.log.  Not part of the product path."""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flair_for_aigle_b200 import native as nv

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/gemm_check.log", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    log.write(s + "\n")
    log.flush()


dev = torch.device("cuda:0")
P(torch.cuda.get_device_name(0))
torch.backends.cuda.matmul.allow_tf32 = False
ok = True
for (M, N, K) in [(128, 128, 64), (128, 128, 128), (256, 128, 256), (1024, 512, 512), (4096, 2048, 512)]:
    try:
        torch.manual_seed(1)
        A = (torch.randn(M, K, device=dev) * 0.5).to(nv.op_dtype())
        B = (torch.randn(N, K, device=dev) / K ** 0.5).to(nv.op_dtype())
        bias = torch.zeros(N, device=dev)
        ref = A.float() @ B.float().t()
        t0 = time.time()
        out = nv.gemm_bf16(A, B, nv.EPI_F32, bias=bias)
        torch.cuda.synchronize()
        err = (out - ref).abs()
        P(f"M{M} N{N} K{K}: max err {err.max().item():.3e} ref max {ref.abs().max().item():.3f} t={time.time()-t0:.3f}s")
        if not (err.max().item() < 1e-2):
            ok = False
            e = err.view(M // 8, 8, N // 16, 16).amax(dim=(1, 3))
            P("err by (8-row group, 16-col group), first 16x8:\n", e[:16, :8].cpu().numpy().round(3))
            sim = nv.gemm_bf16(A, B, nv.EPI_F32, bias=bias, impl="simt")
            torch.cuda.synchronize()
            P("simt max err", (sim - ref).abs().max().item())
            P("out[0,:8]", out[0, :8].tolist(), "ref[0,:8]", ref[0, :8].tolist())
            # K-slice probes: which 16-wide K slices contribute?
            for ks in range(0, min(K, 128), 16):
                A2 = torch.zeros_like(A)
                A2[:, ks:ks + 16] = A[:, ks:ks + 16]
                o2 = nv.gemm_bf16(A2, B, nv.EPI_F32, bias=bias)
                torch.cuda.synchronize()
                r2 = A2.float() @ B.float().t()
                P(f"  k-slice {ks}: err {(o2 - r2).abs().max().item():.3e}  |out| {o2.abs().max().item():.3f} |ref| {r2.abs().max().item():.3f}")
            break
    except Exception as ex:
        ok = False
        P("EXC", M, N, K, repr(ex))
        traceback.print_exc(file=log)
        break


def bench(M, N, K, mode, rps, iters=20):
    A = (torch.randn(M, K, device=dev) * 0.5).to(nv.op_dtype())
    B = (torch.randn(N, K, device=dev) / K ** 0.5).to(nv.op_dtype())
    bias = torch.zeros(N, device=dev)
    resid = torch.zeros(M, N, device=dev) if mode == nv.EPI_RESID_F32 else None
    sumsq = torch.zeros(M // 128, N, device=dev) if mode == nv.EPI_GELU_SUMSQ else None
    out = nv.gemm_bf16(A, B, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps)
    for _ in range(3):
        nv.gemm_bf16(A, B, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        nv.gemm_bf16(A, B, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    C = torch.empty(M, N, device=dev, dtype=nv.op_dtype())
    for _ in range(3):
        torch.matmul(A, B.t(), out=C)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(A, B.t(), out=C)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    P(f"bench M{M} N{N} K{K} mode{mode}: {ms:.4f} ms {2.0*M*N*K/ms/1e9:.1f} TFLOP/s | "
      f"cublas(plain) {ms2:.4f} ms {2.0*M*N*K/ms2/1e9:.1f} TFLOP/s")


if ok:
    try:
        B_ = 16
        bench(B_ * 1024, 2048, 512, nv.EPI_GELU_SUMSQ, 1024)
        bench(B_ * 1024, 512, 2048, nv.EPI_RESID_F32, 1024)
        bench(B_ * 16384, 512, 128, nv.EPI_GELU_SUMSQ, 16384)
        bench(B_ * 16384, 128, 512, nv.EPI_RESID_F32, 16384)
        bench(B_ * 4096, 1024, 256, nv.EPI_GELU_SUMSQ, 4096)
        bench(B_ * 4096, 256, 1024, nv.EPI_RESID_F32, 4096)
        bench(B_ * 256, 4096, 1024, nv.EPI_GELU_SUMSQ, 256)
        bench(B_ * 256, 1024, 4096, nv.EPI_RESID_F32, 256)
        bench(8192, 8192, 8192, nv.EPI_BF16, 8192, iters=5)
    except Exception as ex:
        P("EXC bench", repr(ex))
        traceback.print_exc(file=log)
P("GEMM_CHECK", "OK" if ok else "FAILED")

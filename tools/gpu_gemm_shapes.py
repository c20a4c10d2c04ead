"""GEMM micro-benchmark at the stage shapes for B=37 (and env-selected variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "37"))
def bench(M, N, K, mode, rps, iters=10):
    A = (torch.randn(M, K, device=dev) * 0.5).to(nv.op_dtype())
    Bw = (torch.randn(N, K, device=dev) / K ** 0.5).to(nv.op_dtype())
    bias = torch.zeros(N, device=dev)
    resid = torch.zeros(M, N, device=dev) if mode == nv.EPI_RESID_F32 else None
    sumsq = torch.zeros(M // 128, N, device=dev) if mode == nv.EPI_GELU_SUMSQ else None
    out = nv.gemm_bf16(A, Bw, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps)
    for _ in range(3):
        nv.gemm_bf16(A, Bw, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        nv.gemm_bf16(A, Bw, mode, bias=bias, resid=resid, sumsq=sumsq, rows_per_sample=rps, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"M{M} N{N} K{K} mode{mode}: {ms*1e3:8.1f} us {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s")
print("env", {k: v for k, v in os.environ.items() if k.startswith("FZ_")})
bench(B * 1024, 2048, 512, nv.EPI_GELU_SUMSQ, 1024)
bench(B * 1024, 512, 2048, nv.EPI_RESID_F32, 1024)
bench(B * 256, 4096, 1024, nv.EPI_GELU_SUMSQ, 256)
bench(B * 256, 1024, 4096, nv.EPI_RESID_F32, 256)
bench(B * 1024, 2048, 512, nv.EPI_BF16, 1024)
bench(B * 1024, 512, 2048, nv.EPI_BF16, 1024)

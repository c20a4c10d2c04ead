"""fc1 of ConvNeXt-V2 stage 0 / 1 at B = 37 (K = 128 / 256: epilogue-bound) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
B = 37
for (M, N, K, rps) in ((B * 16384, 512, 128, 16384), (B * 4096, 1024, 256, 4096)):
    A = (torch.randn(M, K, device=dev) * 0.5).to(nv.op_dtype())
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(nv.op_dtype())
    bias = torch.zeros(N, device=dev)
    sq = torch.zeros(M // 128, N, device=dev)
    out = torch.empty(M, N, dtype=nv.op_dtype(), device=dev)
    for _ in range(3):
        nv.gemm_bf16(A, W, nv.EPI_GELU_SUMSQ, bias=bias, sumsq=sq, out=out, rows_per_sample=rps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        nv.gemm_bf16(A, W, nv.EPI_GELU_SUMSQ, bias=bias, sumsq=sq, out=out, rows_per_sample=rps)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    print(f"M{M} N{N} K{K}: {us:8.1f} us  {2.0*M*N*K/us/1e6:6.1f} TFLOP/s  {M*N/us/1e3:6.1f} Gelem/s")

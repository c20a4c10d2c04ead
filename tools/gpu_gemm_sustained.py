"""fc1/fc2 GEMMs of stage 2 at B=37 as the pipeline issues them (per-sample B for fc2, in-place fp32 residual), timed
back to back for long enough to reach the sustained power state.  FZ_GEMM_PAIR=0/1 selects the kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
B = 37
M, C = B * 1024, 512
y = (torch.randn(M, C, device=dev) * 0.5).to(nv.op_dtype())
w1 = (torch.randn(4 * C, C, device=dev) / C ** 0.5).to(nv.op_dtype())
w2 = (torch.randn(B, C, 4 * C, device=dev) / (4 * C) ** 0.5).to(nv.op_dtype())
w2s = w2[0].contiguous()
b1, b2 = torch.zeros(4 * C, device=dev), torch.zeros(C, device=dev)
h = torch.empty(M, 4 * C, dtype=nv.op_dtype(), device=dev)
x = torch.zeros(M, C, device=dev)
sumsq = torch.zeros(M // 128, 4 * C, device=dev)
iters = int(os.environ.get("ITERS", "300"))
def run(name, fn):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name:34s} {us:8.1f} us  {2.0*M*C*4*C/us/1e6:7.1f} TFLOP/s")
print("FZ_GEMM_PAIR =", os.environ.get("FZ_GEMM_PAIR"))
run("fc1 gelu+sumsq", lambda: nv.gemm_bf16(y, w1, nv.EPI_GELU_SUMSQ, bias=b1, sumsq=sumsq, out=h, rows_per_sample=1024))
run("fc2 resid, shared B", lambda: nv.gemm_bf16(h, w2s, nv.EPI_RESID_F32, bias=b2, resid=x, out=x, rows_per_sample=1024))
run("fc2 resid, per-sample B", lambda: nv.gemm_bf16(h, w2, nv.EPI_RESID_F32, bias=b2, resid=x, out=x, rows_per_sample=1024))
def pair():
    nv.gemm_bf16(y, w1, nv.EPI_GELU_SUMSQ, bias=b1, sumsq=sumsq, out=h, rows_per_sample=1024)
    nv.gemm_bf16(h, w2, nv.EPI_RESID_F32, bias=b2, resid=x, out=x, rows_per_sample=1024)
run("fc1+fc2 alternating (per 2 GEMMs)", pair)
def pair_rev():
    nv.gemm_bf16(y, w1, nv.EPI_GELU_SUMSQ, bias=b1, sumsq=sumsq, out=h, rows_per_sample=1024)
    nv.gemm_bf16(h, w2, nv.EPI_RESID_F32 | nv.EPI_REVERSE_TILES, bias=b2, resid=x, out=x, rows_per_sample=1024)
run("same, fc2 walks tiles backwards", pair_rev)

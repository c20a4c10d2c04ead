"""Tiny driver for ncu captures: runs a handful of launches of one kernel family at the bench's
stage-2 shapes (B=16).  usage: python tools/gpu_ncu_targets.py {dwconv|gemm|conv|all}"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flair_for_aigle_b200 import native as nv
what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(os.environ.get("B", "37"))
if what in ("dwconv", "all"):
    for (C, H) in ((512, 32), (128, 128)):
        x = torch.randn(B, H, H, C, device=dev)
        w = torch.randn(49, C, device=dev) * 0.1
        b = torch.randn(C, device=dev); g = torch.rand(C, device=dev) + 0.5; be = torch.randn(C, device=dev)
        out = torch.empty(B, H, H, C, dtype=nv.op_dtype(), device=dev)
        for _ in range(3):
            nv.dwconv7_ln(x, w, b, g, be, out)
if what in ("gemm", "all"):
    for (M, N, K, mode, rps) in ((B * 1024, 2048, 512, nv.EPI_GELU_SUMSQ, 1024), (B * 1024, 512, 2048, nv.EPI_RESID_F32, 1024),
                                 (B * 16384, 512, 128, nv.EPI_GELU_SUMSQ, 16384)):
        A = torch.randn(M, K, device=dev).to(nv.op_dtype()); Bw = (torch.randn(N, K, device=dev) / K ** 0.5).to(nv.op_dtype())
        bias = torch.zeros(N, device=dev)
        resid = torch.zeros(M, N, device=dev) if mode == nv.EPI_RESID_F32 else None
        sq = torch.zeros(M // 128, N, device=dev) if mode == nv.EPI_GELU_SUMSQ else None
        out = None
        for _ in range(3):
            out = nv.gemm_bf16(A, Bw, mode, bias=bias, resid=resid, sumsq=sq, rows_per_sample=rps, out=out)
if what in ("conv", "all"):
    for (H, Cin, Cout) in ((512, 16, 16), (64, 512, 128)):
        x = torch.randn(B, H, H, Cin, device=dev).to(nv.op_dtype())
        w = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).to(nv.op_dtype())
        s = torch.ones(Cout, device=dev); bb = torch.zeros(Cout, device=dev)
        out = torch.empty(B, H, H, Cout, dtype=nv.op_dtype(), device=dev)
        for _ in range(3):
            nv.conv3x3(x, w, s, bb, nv.CONV_RELU_BF16, out=out)
torch.cuda.synchronize()
print("done", what)

"""In-kernel timeline of the persistent GEMM (CTA 0): where do the cycles of a tile go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
names = ["prod_start", "mma_arrive", "acc_free", "first_kb", "mma_issued", "epi_wait", "acc_done", "epi_done"]
for (M, N, K, mode, rps) in ((16384, 2048, 512, nv.EPI_GELU_SUMSQ, 1024), (16384, 512, 2048, nv.EPI_RESID_F32, 1024),
                             (262144, 512, 128, nv.EPI_GELU_SUMSQ, 16384), (8192, 8192, 8192, nv.EPI_BF16, 8192)):
    A = torch.randn(M, K, device=dev).to(nv.op_dtype()); Bw = (torch.randn(N, K, device=dev) / K ** 0.5).to(nv.op_dtype())
    bias = torch.zeros(N, device=dev)
    resid = torch.zeros(M, N, device=dev) if mode == nv.EPI_RESID_F32 else None
    sq = torch.zeros(M // 128, N, device=dev) if mode == nv.EPI_GELU_SUMSQ else None
    out = nv.gemm_bf16(A, Bw, mode, bias=bias, resid=resid, sumsq=sq, rows_per_sample=rps)
    out = nv.gemm_bf16(A, Bw, mode, bias=bias, resid=resid, sumsq=sq, rows_per_sample=rps, out=out)
    tr = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
    nv.lib().fz_gemm_set_trace(ctypes.c_void_p(tr.data_ptr()))
    nv.gemm_bf16(A, Bw, mode, bias=bias, resid=resid, sumsq=sq, rows_per_sample=rps, out=out)
    torch.cuda.synchronize()
    nv.lib().fz_gemm_set_trace(None)
    t = tr.view(64, 8).cpu()
    t0 = t[0, 0].item()
    print(f"=== M{M} N{N} K{K} mode{mode}  (cycles relative to CTA0 producer start)")
    print("tile " + " ".join(f"{n:>11s}" for n in names))
    for i in range(10):
        if t[i, 0].item() == 0: break
        print(f"{i:4d} " + " ".join(f"{(v.item()-t0):11d}" for v in t[i]))

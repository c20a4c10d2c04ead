"""MN-major (transposed operands in place) split-K GEMM against torch: out [M,N] = At^T Bt, At [K,M], Bt [K,N].
Also a one-hot probe that shows where a single (k, m) x (k, n) product lands, to decode a wrong descriptor quickly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
ok = True
for dt in (torch.bfloat16, torch.float16):
    for (M, N, K, splits) in ((256, 128, 64, 1), (256, 128, 128, 1), (512, 128, 1024, 1), (512, 128, 16384, None), (64, 320, 65536, None),
                              (2048, 512, 16384, None), (128, 64, 4096, 7), (16, 64, 262144, None), (1024, 256, 8192, 64), (72, 192, 640, 1)):
        At = (torch.randn(K, M, device=dev) * 0.5).to(dt)
        Bt = (torch.randn(K, N, device=dev) * 0.5).to(dt)
        out = nv.gemm_splitk_tn(At, Bt, splits=splits)
        torch.cuda.synchronize()
        ref = At.double().t() @ Bt.double()
        err = (out.double() - ref).abs().max().item() / ref.std().item()
        good = err < 2e-3
        ok &= good
        print(f"{str(dt)[6:]:9s} M={M:5d} N={N:4d} K={K:7d} splits={splits}: max err / std = {err:.3e} {'ok' if good else 'WRONG'}")
        if not good and K <= 128:
            for (k0, m0, n0) in ((0, 0, 0), (0, 1, 0), (0, 8, 0), (0, 64, 0), (0, 128, 0), (1, 0, 0), (8, 0, 0), (16, 0, 0), (0, 0, 1), (0, 0, 8), (0, 0, 64), (17, 70, 65)):
                a = torch.zeros(K, M, device=dev, dtype=dt); b = torch.zeros(K, N, device=dev, dtype=dt)
                a[k0, m0] = 1; b[k0, n0] = 1
                o = nv.gemm_splitk_tn(a, b, splits=1)
                torch.cuda.synchronize()
                nz = torch.nonzero(o).tolist()
                print(f"   one-hot k={k0} m={m0} n={n0} -> nonzeros at {nz[:6]} values {[float(o[i, j]) for i, j in nz[:6]]}")
print("MN_MAJOR_OK" if ok else "MN_MAJOR_FAIL")

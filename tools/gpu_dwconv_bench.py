"""Depthwise 7x7 + LayerNorm at the four ConvNeXt-V2-base stage shapes (B = 37): FZ_DWCONV_PERSIST=0/1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
B = 37
print("FZ_DWCONV_PERSIST =", os.environ.get("FZ_DWCONV_PERSIST"))
for C, H in ((128, 128), (256, 64), (512, 32), (1024, 16)):
    x = torch.randn(B, H, H, C, device=dev)
    w = torch.randn(49, C, device=dev) * 0.1
    b = torch.randn(C, device=dev); g = torch.rand(C, device=dev) + 0.5; be = torch.randn(C, device=dev)
    out = torch.empty(B, H, H, C, dtype=nv.op_dtype(), device=dev)
    for _ in range(3):
        nv.dwconv7_ln(x, w, b, g, be, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        nv.dwconv7_ln(x, w, b, g, be, out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"C={C:5d} H={H:4d}: {us:8.1f} us   {2.0*49*B*H*H*C/us/1e6:6.2f} TFLOP/s fp32")

"""The HBM-bound decoder tail at B = 37: conv 16->16 @512^2 and the head conv 16->19 with the argmax epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
B, H = 37, 512
x = torch.randn(B, H, H, 16, device=dev).to(nv.op_dtype())
w = (torch.randn(16, 3, 3, 16, device=dev) / 12).to(nv.op_dtype())
wh = torch.zeros(32, 3, 3, 16, device=dev).to(nv.op_dtype()); wh[:19] = (torch.randn(19, 3, 3, 16, device=dev) / 12).to(nv.op_dtype())
s = torch.ones(32, device=dev); b = torch.zeros(32, device=dev)
out = torch.empty(B, H, H, 16, dtype=nv.op_dtype(), device=dev)
plan = torch.zeros(B, 6, dtype=torch.int32, device=dev); plan[:, 4:] = 384
for i in range(B):
    plan[i, 2] = (i // 8) * 384; plan[i, 3] = (i % 8) * 384
raster = torch.zeros(5 * 384, 8 * 384, dtype=torch.uint8, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("conv 16->16 relu      %.1f us" % t(lambda: nv.conv3x3(x, w, s[:16], b[:16], nv.CONV_RELU_BF16, out=out)))
print("head 16->19 argmax    %.1f us" % t(lambda: nv.conv3x3(x, wh, None, b, nv.CONV_ARGMAX_RASTER, cout=19, plan=plan, own=None, raster=raster, margin=64)))

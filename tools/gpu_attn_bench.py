"""Window-attention micro-benchmark at the four Swin-base stage shapes (B=37 tiles of 512^2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "37"))
only = os.environ.get("STAGE")
for i, (H, heads) in enumerate([(128, 4), (64, 8), (32, 16), (16, 32)]):
    if only is not None and int(only) != i:
        continue
    C = heads * 32
    qkv = (torch.randn(B, H, H, 3 * C, device=dev)).to(nv.op_dtype())
    bias = torch.randn(3 * C, device=dev).to(nv.op_dtype())
    table = torch.randn(heads, 529, device=dev) * 0.5
    out = torch.empty(B, H, H, C, dtype=nv.op_dtype(), device=dev)
    for shift in (0, 6):
        for _ in range(2):
            nv.swin_window_attn(qkv, bias, table, out, heads, 12, shift, 32 ** -0.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            nv.swin_window_attn(qkv, bias, table, out, heads, 12, shift, 32 ** -0.5)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 5 * 1e3
        nwin = ((H + 11) // 12) ** 2
        fl = B * nwin * heads * 4.0 * 144 * 144 * 32
        print(f"stage {i} H={H} heads={heads} shift={shift}: {us:8.1f} us  {fl/us/1e6:6.1f} TFLOP/s  ({B*nwin*heads} CTAs)")

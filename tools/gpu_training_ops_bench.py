"""HBM roofline of the training step's loss / optimizer kernels at BASELINE.json configs[4] sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
from flair_for_aigle_b200 import native as nv
dev = torch.device("cuda:0")
peak = 6529.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
B, C, H, W = 16, 19, 512, 512
logits = torch.randn(B, C, H, W, device=dev)
tg = torch.randint(0, C, (B, H, W), device=dev, dtype=torch.int32)
w = torch.ones(C, device=dev); w[15:] = 0
out, lse, preds = nv.ce_loss_forward(logits, tg, w, 1.0)
dl = torch.empty_like(logits)
px = B * H * W
ws = torch.empty(int(nv.lib().fz_ce_workspace_doubles(B, H, W)), dtype=torch.float64, device=dev)
L, P, S = nv.lib(), nv._ptr, nv._stream
# the C entry point with preallocated buffers (the Python wrapper's allocations would dominate a 60 us kernel)
us = t(lambda: L.fz_ce_loss_forward(P(logits), P(tg), P(w), 1.0, P(lse), P(preds), P(ws), P(out), B, C, H, W, S()))
by = px * (C * 4 + 4 + 4 + 4)
print(f"ce_loss_forward  (16,19,512,512): {us:7.1f} us  {by / us / 1e3:7.0f} GB/s = {by / us / 1e3 / peak:.0%} of {peak:.0f}")
us = t(lambda: L.fz_ce_loss_backward(P(logits), P(tg), P(w), 1.0, P(lse), P(out), 1.0, P(dl), B, C, H, W, S()))
by = px * (2 * C * 4 + 4 + 4)
print(f"ce_loss_backward (16,19,512,512): {us:7.1f} us  {by / us / 1e3:7.0f} GB/s = {by / us / 1e3 / peak:.0%}")
n = 183_000_000
p, g, m, v = (torch.randn(n, device=dev) for _ in range(4)); v.abs_()
us = t(lambda: nv.adamw_step(p, g, m, v, 5e-5, 0.9, 0.999, 1e-8, 0.01, 3))
by = n * 28
print(f"adamw_step (183 M parameters):     {us:7.1f} us  {by / us / 1e3:7.0f} GB/s = {by / us / 1e3 / peak:.0%}")

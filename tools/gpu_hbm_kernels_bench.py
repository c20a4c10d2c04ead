"""Achieved HBM bandwidth of the memory-bound kernels either side of the model, at BASELINE.json configs[1]
sizes: one batch of 37 tiles of 512^2, margin 64, 19 classes, on a 10 000 x 10 000 zone.

  gather_u8 / gather_f32     the tile feeder (dataset.py:89-124): raster window -> tile
  crop_argmax_write          crop + argmax + ownership write (inference.py:295-352), fp16 NHWC (cstride 24), fp32 NCHW
  crop_softmax_write         class_prob output: round(softmax * 255) -> 19 uint8 planes
  crop_softmax_accumulate    the accumulating variant (inference.py:468-564): fp32 canvas +=
  canvas_argmax              logits_to_labels_and_confidence (inference.py:566-572) over the 19 x 10k x 10k canvas
  confusion_matrix           the metrics' label x prediction counts (tasks_module.py:212,274; prediction_writer.py:64)

Algorithmic bytes = what the kernel must read + write once (stated per row below); time = CUDA events over 20 launches after
3 warm-ups, the batch's tensors (>= 370 MB) far exceed nothing but L2 for the small ones, so a 256 MB buffer is written between
launches to flush L2.  Peak = MEASURED_PEAKS.json hbm_gbs (copy bandwidth).  Writes gpurun_out/r2_hbm_kernels.txt."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from flair_for_aigle_b200 import native as nv  # noqa: E402

dev = torch.device("cuda:0")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    peak_src = "MEASURED_PEAKS.json"
except Exception:
    PEAK, peak_src = 6650.0, "B200_PROFILING.md fallback"
B, P, M, C, ZONE = 37, 512, 64, 19, 10000
S = P - 2 * M
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1)                                  # evict the previous launch's lines from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3                                # us


# 37 consecutive tiles of the zone's grid (stride 384): disjoint inner windows
plan = np.zeros((B, 6), np.int32)
for i in range(B):
    r, c = divmod(i, 25)
    plan[i] = (r * S - M, c * S - M, r * S, c * S, S, S)
plan_d, own_d = torch.from_numpy(plan).to(dev), None      # disjoint write windows: no ownership clipping needed
raster_u8 = torch.randint(0, 255, (4, ZONE, ZONE), dtype=torch.uint8, device=dev)
raster_f32 = torch.randn(1, ZONE, ZONE, device=dev)
origins = plan_d[:, :2].contiguous()
out_raster = torch.zeros((ZONE, ZONE), dtype=torch.uint8, device=dev)
out_planes = torch.zeros((C, ZONE, ZONE), dtype=torch.uint8, device=dev)
canvas = torch.zeros((C, ZONE, ZONE), dtype=torch.float32, device=dev)
op = nv.op_dtype()
l16 = (torch.randn(B, P, P, 24, device=dev) * 3).to(op)
l32 = torch.randn(B, C, P, P, device=dev) * 3
l32h = (torch.randn(B, P, P, 20, device=dev) * 3)
tiles_u8 = torch.empty((B, P, P, 4), dtype=torch.uint8, device=dev)
tiles_f32 = torch.empty((B, 4, P, P), dtype=torch.float32, device=dev)
tiles_dem = torch.empty((B, 1, P, P), dtype=torch.float32, device=dev)
mean4, std4 = torch.zeros(4, device=dev), torch.ones(4, device=dev)
px_in, px_out = B * P * P, B * S * S
# labels and predictions for the metric kernel: runs of one class, like a class raster (16 x 512^2 = one training batch; and
# a 10k^2 zone's worth, what an evaluation over a predicted raster counts)
lab16 = torch.randint(0, C, (16 * P * P // 64,), device=dev, dtype=torch.int32).repeat_interleave(64)
prd16 = torch.where(torch.rand(lab16.shape, device=dev) < 0.8, lab16, torch.randint(0, C, lab16.shape, device=dev, dtype=torch.int32))
labz = torch.randint(0, C, (ZONE * ZONE // 64,), device=dev, dtype=torch.int32).repeat_interleave(64)
prdz = torch.where(torch.rand(labz.shape, device=dev) < 0.8, labz, torch.randint(0, C, labz.shape, device=dev, dtype=torch.int32))
cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
rows = [
    ("confusion_matrix 16 x 512^2 int32 labels + predictions (8 B / px)", lambda: nv.confusion_matrix(lab16, prd16, C, out=cm), lab16.numel() * 8),
    ("confusion_matrix 10k x 10k (8 B / px)", lambda: nv.confusion_matrix(labz, prdz, C, out=cm), labz.numel() * 8),
    ("gather_u8 (4 B read + 4 B write / input px)", lambda: nv.gather_tiles_u8(raster_u8, origins, P, out=tiles_u8), px_in * 8),
    ("gather_f32 from uint8 (4 B read + 16 B write / px)", lambda: nv.gather_tiles_f32(raster_u8, origins, P, mean4, std4, out=tiles_f32), px_in * 20),
    ("gather_f32 from float32 DEM (4 B + 4 B / px)", lambda: nv.gather_tiles_f32(raster_f32, origins, P, mean4[:1], std4[:1], out=tiles_dem), px_in * 8),
    (f"crop_argmax_write {str(op)[6:]} NHWC cstride 24 (read 48 B of the row's 24 x 2, write 1 B / output px)",
     lambda: nv.crop_argmax_write(l16, nv.NHWC, M, plan_d, own_d, out_raster, n_cls=C), px_out * 49),
    ("crop_argmax_write fp32 NHWC cstride 20 (80 B + 1 B / output px)",
     lambda: nv.crop_argmax_write(l32h, nv.NHWC, M, plan_d, own_d, out_raster, n_cls=C), px_out * 81),
    ("crop_argmax_write fp32 NCHW (76 B + 1 B / output px)",
     lambda: nv.crop_argmax_write(l32, nv.NCHW, M, plan_d, own_d, out_raster), px_out * 77),
    ("crop_softmax_write fp32 NCHW -> 19 uint8 planes (76 B + 19 B / output px)",
     lambda: nv.crop_softmax_write(l32, nv.NCHW, M, plan_d, own_d, out_planes), px_out * 95),
    ("crop_softmax_accumulate fp32 NCHW -> fp32 canvas (76 B + 76 B read + 76 B write / output px)",
     lambda: nv.crop_softmax_accumulate(l32, nv.NCHW, M, plan_d, None, canvas, plan_host=plan), px_out * 228),
    ("crop_softmax_accumulate, one launch per tile (no host plan: the round-1 behaviour)",
     lambda: nv.crop_softmax_accumulate(l32, nv.NCHW, M, plan_d, None, canvas), px_out * 228),
    ("canvas_argmax 19 x 10k x 10k fp32 -> uint8 + fp32 confidence (76 B + 5 B / px)",
     lambda: nv.canvas_argmax(canvas, want_confidence=True), ZONE * ZONE * 81),
]
lines = [f"HBM-bound kernels at configs[1] sizes (batch {B} x {P}^2, margin {M}, {C} classes, zone {ZONE}^2); peak {PEAK} GB/s ({peak_src}); "
         f"L2 flushed between launches", f"{'kernel (algorithmic bytes)':100s} {'us':>9s} {'GB/s':>8s} {'frac':>6s}"]
for name, fn, nbytes in rows:
    us = timed(fn, n=5 if "canvas_argmax" in name else 20)
    gbs = nbytes / us / 1e3
    lines.append(f"{name:100s} {us:9.1f} {gbs:8.0f} {gbs / PEAK:6.3f}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/r2_hbm_kernels.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

"""Polygonisation stage on a 10 000 x 10 000 class raster: GPU labelling (CUDA events) + host tracing vs the oracle's
scipy.ndimage.label per class (the structure rasterio.features.shapes walks per class mask in the reference)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flair_for_aigle_b200 import native as nv
from flair_for_aigle_b200.synthetic import synthetic_raster
from flair_for_aigle_b200.flair_zonal_detection.polygonize import raster_to_polygons

N = int(os.environ.get("N", "10000"))
arr = synthetic_raster(N, N, seed=2025)
raster = ((arr[0].astype(np.int32) + arr[1]) // 40 % 19).astype(np.uint8)
dev = torch.device("cuda:0")
t = torch.from_numpy(raster).to(dev)
for _ in range(2):
    labels = nv.ccl_label(t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    labels = nv.ccl_label(t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
px = N * N
print(f"ccl_label {N}x{N}: {ms:.2f} ms = {px / ms / 1e6:.2f} Gpx/s; algorithmic 20 B/px -> {20 * px / ms / 1e6:.0f} GB/s")
t0 = time.time(); roots, areas, classes = nv.ccl_components(t, labels, min_area_px=25, ignore_class=18); torch.cuda.synchronize(); t1 = time.time()
print(f"components: {nv.ccl_components.last_total} total, {roots.size} with >= 25 px and class != 18 ({(t1 - t0) * 1e3:.1f} ms incl. table download)")
t0 = time.time(); table = raster_to_polygons((t, 700000.0, 6600000.0, 0.2), device=dev); t1 = time.time()
nr = len(table.geometry.ring_perm)
print(f"raster_to_polygons (label + table + D2H + trace + simplify): {t1 - t0:.2f} s, {len(table)} polygons >= 1 m^2, {nr} rings, {table.geometry.xy.shape[0]} points")
smooth = ((arr[0].astype(np.int32) // 32) % 19).astype(np.uint8)          # large regions, little speckle
ts = torch.from_numpy(smooth).to(dev)
nv.ccl_label(ts); torch.cuda.synchronize()
e0.record(); ls = nv.ccl_label(ts); e1.record(); torch.cuda.synchronize()
print(f"smooth map: ccl_label {e0.elapsed_time(e1):.2f} ms")
t0 = time.time(); table = raster_to_polygons((ts, 700000.0, 6600000.0, 0.2), device=dev); t1 = time.time()
print(f"smooth map: raster_to_polygons {t1 - t0:.2f} s, {len(table)} polygons, {len(table.geometry.ring_perm)} rings, {table.geometry.xy.shape[0]} points")
if os.environ.get("CPU", "1") == "1":
    from scipy import ndimage
    from oracle.polygons import FOUR
    sub = raster[: N // 4]                       # bounded sample: a quarter of the rows
    t0 = time.time()
    for cls in np.unique(sub):
        ndimage.label(sub == cls, structure=FOUR)
    t1 = time.time()
    print(f"cpu scipy.ndimage.label per class on {sub.shape[0]}x{sub.shape[1]}: {t1 - t0:.2f} s -> {sub.size / (t1 - t0) / 1e6:.1f} Mpx/s (1 core)")

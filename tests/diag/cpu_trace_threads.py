"""Host ring tracer: thread scaling on this machine's cores (no GPU needed)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flair_for_aigle_b200 import native as nv
from flair_for_aigle_b200.synthetic import synthetic_raster
from oracle.polygons import component_table, label_components
N = int(os.environ.get("N", "3000"))
arr = synthetic_raster(N, N, seed=2025)
raster = ((arr[0].astype(np.int32) + arr[1]) // 40 % 19).astype(np.uint8)
labels = label_components(raster)
roots, areas, classes = component_table(raster, labels)
keep = roots[(classes != 18) & (areas * 0.04 >= 1.0)]
print("cpus", os.cpu_count(), "raster", N, "kept components", keep.size)
for t in ("1", "2", "4", "8", "16"):
    os.environ["FZ_TRACE_THREADS"] = t
    t0 = time.time(); r = nv.trace_rings(labels, keep, 0.5)
    print(f"{t:>2s} threads: {time.time() - t0:.3f} s, {r[0].size} rings")

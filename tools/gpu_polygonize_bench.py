"""raster_to_polygons on a synthetic class raster: GPU labelling + host ring tracing, by tracer thread
count.  python tools/gpu_polygonize_bench.py [--size 5000] [--out profiles/r2_polygonize_bench.txt]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=5000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    from flair_for_aigle_b200.flair_zonal_detection.polygonize import raster_to_polygons
    n = args.size
    rng = np.random.default_rng(0)
    cells = rng.integers(0, 19, (n // 40 + 2, n // 40 + 2)).astype(np.uint8)
    raster = np.kron(cells, np.ones((40, 40), np.uint8))[:n, :n].copy()
    raster[rng.random(raster.shape) < 0.02] = 7                     # speckle: millions of tiny components and holes
    dev = torch.device("cuda:0")
    src = (torch.from_numpy(raster).to(dev), 700000.0, 6600000.0, 0.2, "EPSG:2154")
    lines = [f"raster_to_polygons, {n} x {n} synthetic class raster (40 px cells of 19 classes + 2 % speckle), min_area 0.1 m2, "
             f"simplification 0.1 m; host: {os.cpu_count()} cores"]
    raster_to_polygons(src, min_area=0.1, simplification=0.1, device=dev)          # warm-up
    ref = None
    for threads in (1, 2, 4, 8, os.cpu_count()):
        os.environ["FZ_TRACE_THREADS"] = str(threads)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        table = raster_to_polygons(src, min_area=0.1, simplification=0.1, device=dev)
        dt = time.perf_counter() - t0
        npts = int(table.geometry.ring_off[-1])
        sig = (len(table), npts, float(table.area.sum()), float(table.geometry.xy.sum()))
        same = ref is None or sig == ref
        ref = ref or sig
        lines.append(f"  tracer threads {threads:3d}: whole call {dt * 1e3:8.1f} ms   {len(table)} polygons, {npts} points, "
                     f"same result as 1 thread: {same}")
    os.environ.pop("FZ_TRACE_THREADS", None)
    print("\n".join(lines), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

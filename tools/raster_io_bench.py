"""Host benchmark of the raster file I/O at the two ends of the zonal path: libfz_rasterio.so
(block-parallel, all cores) next to libtiff through Pillow (one core -- what the round-1 code used, and what GDAL does for
the reference's per-window reads / LZW window writes).  No GPU needed.

    python tools/raster_io_bench.py [--size 10000] [--out profiles/r2_raster_io_bench.txt]

Rasters: the synthetic 4-band uint8 zone of bench.py (configs[1] size by default) as a tiled, pixel-interleaved LZW GeoTIFF
with predictor 2 (GDAL's usual ortho layout), and a blocky 19-class argmax raster of the same size."""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from PIL import Image  # noqa: E402

from flair_for_aigle_b200 import raster_io as rio  # noqa: E402

Image.MAX_IMAGE_PIXELS = None


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=10000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    n = args.size
    lines = []

    def say(s):
        print(s, flush=True)
        lines.append(s)

    from flair_for_aigle_b200.synthetic import synthetic_raster
    ortho = np.ascontiguousarray(synthetic_raster(n, n, 4, seed=2025))
    rng = np.random.default_rng(0)
    cells = rng.integers(0, 19, (n // 24 + 2, n // 24 + 2)).astype(np.uint8)
    classes = np.ascontiguousarray(np.kron(cells, np.ones((24, 24), np.uint8))[:n, :n])[None]
    cores = os.cpu_count()
    say(f"raster I/O, {n} x {n} px, host with {cores} cores; times = best of {args.reps}; MB = raw pixel bytes / 1e6")
    say("")
    with tempfile.TemporaryDirectory() as tmp:
        # ---------------------------------------------------------------- output side: the argmax raster (100 MB at 10k^2)
        mb = classes.nbytes / 1e6
        p_lib, p_1, p_n = (os.path.join(tmp, x) for x in ("cls_libtiff.tif", "cls_1.tif", "cls_n.tif"))
        t_lib = best(lambda: Image.fromarray(classes[0]).save(p_lib, format="TIFF", compression="tiff_lzw"), args.reps)
        t_1 = best(lambda: rio.write_geotiff(p_1, classes, 7e5, 6.6e6, 0.2, epsg=2154, threads=1), args.reps)
        t_n = best(lambda: rio.write_geotiff(p_n, classes, 7e5, 6.6e6, 0.2, epsg=2154), args.reps)
        say(f"WRITE argmax raster ({mb:.0f} MB, LZW like the reference's profile)")
        say(f"  libtiff via Pillow, strips, 1 thread        {t_lib * 1e3:8.1f} ms  {mb / t_lib:8.0f} MB/s  file {os.path.getsize(p_lib) / 1e6:6.1f} MB")
        say(f"  libfz_rasterio, 512 tiles, 1 thread         {t_1 * 1e3:8.1f} ms  {mb / t_1:8.0f} MB/s  file {os.path.getsize(p_1) / 1e6:6.1f} MB")
        say(f"  libfz_rasterio, 512 tiles, {cores:2d} threads        {t_n * 1e3:8.1f} ms  {mb / t_n:8.0f} MB/s  ({t_lib / t_n:.1f}x libtiff)")
        p_cog = os.path.join(tmp, "cls_cog.tif")
        t_cog = best(lambda: rio.write_geotiff(p_cog, classes, 7e5, 6.6e6, 0.2, epsg=2154, cog=True, overviews=-1), args.reps)
        say(f"  COG (nearest overviews to <= 512, IFDs first)  {t_cog * 1e3:8.1f} ms  {mb / t_cog:8.0f} MB/s  file {os.path.getsize(p_cog) / 1e6:6.1f} MB, "
            f"{rio.tiff_info(p_cog).overviews} overviews")
        t_conv = best(lambda: rio.convert_to_cog(p_n, os.path.join(tmp, "conv_cog.tif")), args.reps)
        say(f"  convert_to_cog(file -> COG file)             {t_conv * 1e3:8.1f} ms  (postprocess.py:33-52: decode + pyramid + encode)")
        t_rd = best(lambda: rio.read_raster(p_n), args.reps)
        t_rd_lib = best(lambda: np.asarray(Image.open(p_lib)), args.reps)
        say(f"  read it back: libfz_rasterio {t_rd * 1e3:.1f} ms ({mb / t_rd:.0f} MB/s), libtiff {t_rd_lib * 1e3:.1f} ms ({mb / t_rd_lib:.0f} MB/s)")
        got = rio.read_raster(p_n)[0]
        assert np.array_equal(got, classes) and np.array_equal(np.asarray(Image.open(p_n)), classes[0])
        say("")
        # ---------------------------------------------------------------- input side: the 4-band ortho (400 MB at 10k^2)
        mb = ortho.nbytes / 1e6
        p_in = os.path.join(tmp, "ortho.tif")
        t_w = best(lambda: rio.write_geotiff(p_in, ortho, 7e5, 6.6e6, 0.2, epsg=2154, pixel_interleave=True, predictor=2), 1)
        say(f"READ 4-band uint8 ortho ({mb:.0f} MB raw; tiled LZW, pixel-interleaved, predictor 2: file {os.path.getsize(p_in) / 1e6:.1f} MB, "
            f"written in {t_w * 1e3:.0f} ms)")
        dst = np.empty_like(ortho)
        t_1 = best(lambda: rio.read_raster(p_in, out=dst, threads=1), args.reps)
        t_n = best(lambda: rio.read_raster(p_in, out=dst), args.reps)
        assert np.array_equal(dst, ortho)
        p_in3 = os.path.join(tmp, "ortho_rgb.tif")
        rio.write_geotiff(p_in3, ortho[:3], pixel_interleave=True, predictor=2)
        t_lib = best(lambda: np.ascontiguousarray(np.asarray(Image.open(p_in3)).transpose(2, 0, 1)), args.reps)
        say(f"  libtiff via Pillow (RGB only: 3 of the 4 bands), 1 thread, + transpose to (C,H,W)   {t_lib * 1e3:8.1f} ms  {0.75 * mb / t_lib:8.0f} MB/s")
        say(f"  libfz_rasterio, 1 thread, decoded straight to (C,H,W)                              {t_1 * 1e3:8.1f} ms  {mb / t_1:8.0f} MB/s")
        say(f"  libfz_rasterio, {cores:2d} threads                                                         {t_n * 1e3:8.1f} ms  {mb / t_n:8.0f} MB/s")
        # the reference's access pattern: one boundless 512 x 512 window per tile (dataset.py:89-117), 729 tiles at 10k^2
        from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference  # noqa: F401
        stride, P, m = 384, 512, 64
        origins = sorted({min(max(o, -m), n - P + m) for o in range(-m, n, stride)})
        wins = [(r, c) for r in origins for c in origins]
        info = rio.tiff_info(p_in)
        t0 = time.perf_counter()
        for r, c in wins[:200]:
            rio.read_window(p_in, r, c, P, P, info=info, threads=1)
        t_win = (time.perf_counter() - t0) / 200
        say(f"  per-tile windows like the reference ({len(wins)} boundless 512 x 512 reads, 1 thread each): {t_win * 1e3:.2f} ms / window -> "
            f"{t_win * len(wins) * 1e3:.0f} ms per zone vs {t_n * 1e3:.0f} ms for ONE block-parallel pass into the upload buffer")
        # ---------------------------------------------------------------- JPEG 2000 orthos (the reference's product inputs)
        try:
            from flair_for_aigle_b200 import openjpeg as oj
            m = min(n, 4000)
            crop = np.ascontiguousarray(ortho[:, :m, :m].transpose(1, 2, 0))
            say("")
            say(f"READ JPEG 2000 ortho ({m} x {m} x 4, tiles of 1024; OpenJPEG {oj.lib().opj_version().decode()} from Pillow's wheel)")
            for label, kw in (("lossless (reversible 5/3)", dict(irreversible=False)),
                              ("lossy 1:10 (irreversible 9/7)", dict(irreversible=True, quality_mode="rates", quality_layers=[10]))):
                pj = os.path.join(tmp, "ortho.jp2")
                Image.fromarray(crop).save(pj, format="JPEG2000", tile_size=(1024, 1024), **kw)
                mbj = crop.nbytes / 1e6
                t_pil = best(lambda: np.ascontiguousarray(np.asarray(Image.open(pj)).transpose(2, 0, 1)), 1)
                dstj = np.empty((4, m, m), np.uint8)
                t_1 = best(lambda: oj.read_rows(pj, out=dstj, threads=1), 1)
                t_n = best(lambda: oj.read_rows(pj, out=dstj), 2)
                t_strip = best(lambda: oj.read_rows(pj, m // 2, min(m, m // 2 + 1024)), 2)
                say(f"  {label:30s} file {os.path.getsize(pj) / 1e6:6.1f} MB: Pillow (1 thread, whole image) {t_pil * 1e3:7.0f} ms = {mbj / t_pil:5.0f} MB/s | "
                    f"OpenJPEG direct 1 thread {t_1 * 1e3:7.0f} ms | {cores} threads {t_n * 1e3:7.0f} ms = {mbj / t_n:5.0f} MB/s | "
                    f"one 1024-row strip {t_strip * 1e3:6.0f} ms")
        except Exception as ex:  # noqa: BLE001
            say(f"  JPEG 2000 section skipped: {ex!r}")
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

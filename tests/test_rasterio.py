"""CPU: libfz_rasterio.so -- the block-parallel TIFF / BigTIFF / GeoTIFF / COG reader and writer at the two ends of the zonal
path (include/flair_zonal_rasterio.h).

The independent implementation it is pinned against is libtiff, through Pillow, in BOTH directions: files written here are
decoded by libtiff, files written by libtiff are decoded here; on top of that the reader's boundless windows are compared
with numpy slicing of the zero-padded array (what rasterio's ``read(window=..., boundless=True, fill_value=0)`` returns,
flair_zonal_detection/dataset.py:108-115)."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest
from PIL import Image

from flair_for_aigle_b200 import raster_io as rio

Image.MAX_IMAGE_PIXELS = None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEFT, TOP, RES = 700000.0, 6600000.0, 0.2


def class_map(rng, h, w, c=1, cell=16, n_cls=19):
    """blocky uint8 content like an argmax raster"""
    base = rng.integers(0, n_cls, (c, h // cell + 2, w // cell + 2)).astype(np.uint8)
    return np.ascontiguousarray(np.kron(base, np.ones((cell, cell), np.uint8))[:, :h, :w])


def pillow_frames(path):
    out = []
    with Image.open(path) as im:
        for k in range(getattr(im, "n_frames", 1)):
            im.seek(k)
            out.append(np.asarray(im).copy())
    return out


def test_library_exports_every_declared_symbol():
    from flair_for_aigle_b200.build import build_rasterio
    path = build_rasterio()
    text = open(os.path.join(ROOT, "include", "flair_zonal_rasterio.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(fzio_[a-z0-9_]+)\s*\(", text)))
    handle = ctypes.CDLL(str(path))
    assert len(declared) == 9
    for name in declared:
        assert hasattr(handle, name), name
    assert sorted(rio.exported_symbols()) == declared
    assert rio.lib().fzio_abi_version() == rio.ABI_VERSION == 1
    # the ctypes mirrors of the two structs have the C layout
    assert ctypes.sizeof(rio._Info) == 16 + 16 * 4 + 32 and ctypes.sizeof(rio._WriteOpts) == 16 * 4 + 24


@pytest.mark.parametrize("kind", ["random", "zeros", "ramp", "low_entropy", "runs"])
def test_lzw_codec_round_trip(kind):
    rng = np.random.default_rng(1)
    for n in (0, 1, 2, 3, 255, 4096, 70001, 400000):
        a = {"random": rng.integers(0, 256, n), "zeros": np.zeros(n), "ramp": np.arange(n) % 7,
             "low_entropy": rng.integers(0, 3, n), "runs": np.repeat(rng.integers(0, 19, n // 50 + 1), 50)[:n]}[kind]
        a = a.astype(np.uint8).tobytes()
        enc = rio.lzw_encode(a)
        assert rio.lzw_decode(enc, n) == a
        assert len(enc) <= rio.lib().fzio_lzw_bound(n)
    with pytest.raises(rio.RasterIOError):
        rio.lzw_decode(b"\x00\x00\x00\x00", 16)          # does not start with ClearCode


def test_lzw_against_libtiff_streams(tmp_path):
    """One strip = one LZW stream.  libtiff's streams decode here; on noise and flat data this encoder's bytes EQUAL libtiff's
    (same table-full reset, code-width changes and end of stream); on compressible class maps libtiff additionally resets
    its table when its running compression ratio drops (a heuristic, not part of the format), so there only the size is
    compared."""
    rng = np.random.default_rng(2)
    for name, arr in (("noise", rng.integers(0, 256, (300, 500)).astype(np.uint8)), ("classes", class_map(rng, 300, 500)[0]),
                      ("flat", np.full((300, 500), 7, np.uint8))):
        p = str(tmp_path / f"{name}.tif")
        Image.fromarray(arr).save(p, format="TIFF", compression="tiff_lzw")
        with Image.open(p) as im:
            offs, cnts, rps = im.tag_v2[273], im.tag_v2[279], im.tag_v2[278]
        raw = open(p, "rb").read()
        for k, (o, c) in enumerate(zip(offs, cnts)):
            rows = arr[k * rps:(k + 1) * rps]
            assert rio.lzw_decode(raw[o:o + c], rows.size) == rows.tobytes()
            mine = rio.lzw_encode(rows.tobytes())
            if name == "classes":
                assert len(mine) <= 1.05 * c, (name, k, len(mine), c)
            else:
                assert mine == raw[o:o + c], (name, k)


@pytest.mark.parametrize("shape", [(700, 1000), (512, 512), (513, 511), (17, 33), (1, 1)])
def test_written_tiles_decode_in_libtiff_and_round_trip(tmp_path, shape):
    rng = np.random.default_rng(3)
    a = class_map(rng, *shape)
    for comp in ("lzw", "deflate", "none"):
        for pred in ((1, 2) if comp != "none" else (1,)):
            p = str(tmp_path / f"{comp}{pred}.tif")
            rio.write_geotiff(p, a, LEFT, TOP, RES, epsg=2154, compression=comp, predictor=pred, block=256)
            assert np.array_equal(pillow_frames(p)[0], a[0])
            got, info = rio.read_raster(p)
            assert np.array_equal(got, a)
            assert (info.width, info.height, info.count, info.dtype) == (shape[1], shape[0], 1, np.uint8)
            assert info.tiled and info.block_w == 256 and info.predictor == pred and not info.bigtiff
            assert (info.left, info.top, info.res_x, info.res_y, info.epsg, info.crs) == (LEFT, TOP, RES, RES, 2154, "EPSG:2154")
    with pytest.raises(rio.RasterIOError, match="predictor 2 needs"):
        rio.write_geotiff(str(tmp_path / "x.tif"), a, compression="none", predictor=2)


def test_incompressible_data_and_thread_count_invariance(tmp_path):
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, (1, 1030, 1100)).astype(np.uint8)
    p1, p8 = str(tmp_path / "t1.tif"), str(tmp_path / "t8.tif")
    rio.write_geotiff(p1, a, threads=1)
    rio.write_geotiff(p8, a, threads=8)
    assert open(p1, "rb").read() == open(p8, "rb").read()               # the file does not depend on the thread count
    assert np.array_equal(pillow_frames(p1)[0], a[0])
    assert np.array_equal(rio.read_raster(p1, threads=1)[0], a) and np.array_equal(rio.read_raster(p1, threads=5)[0], a)


@pytest.mark.parametrize("mode,c", [("L", 1), ("RGB", 3), ("RGBA", 4)])
def test_libtiff_written_strips_decode_here(tmp_path, mode, c):
    """Pillow / libtiff writes pixel-interleaved strips; the last strip is ragged."""
    rng = np.random.default_rng(5)
    for arr in (class_map(rng, 601, 777, c), rng.integers(0, 256, (c, 601, 777)).astype(np.uint8)):
        for comp in ("tiff_lzw", "tiff_adobe_deflate", None):
            p = str(tmp_path / "p.tif")
            Image.fromarray(arr[0] if c == 1 else arr.transpose(1, 2, 0), mode).save(p, format="TIFF", compression=comp)
            got, info = rio.read_raster(p)
            assert not info.tiled and info.count == c
            assert np.array_equal(got, arr), (mode, comp)
            if c > 1:                                                    # rasterio's indexes=[...] : 1-based, any order
                sel, _ = rio.read_raster(p, bands=[c, 1])
                assert np.array_equal(sel, arr[[c - 1, 0]])


def test_libtiff_predictor_uint16_float32(tmp_path):
    rng = np.random.default_rng(6)
    a8 = class_map(rng, 300, 400, 3)
    p = str(tmp_path / "pred.tif")
    Image.fromarray(a8.transpose(1, 2, 0), "RGB").save(p, format="TIFF", compression="tiff_lzw", tiffinfo={317: 2})
    got, info = rio.read_raster(p)
    assert info.predictor == 2 and np.array_equal(got, a8)
    a16 = rng.integers(0, 65536, (1, 211, 333)).astype(np.uint16)
    p = str(tmp_path / "u16.tif")
    Image.fromarray(a16[0]).save(p, format="TIFF", compression="tiff_adobe_deflate")
    got, info = rio.read_raster(p)
    assert info.dtype == np.uint16 and np.array_equal(got, a16)
    dem = (rng.standard_normal((1, 211, 333)) * 50 + 400).astype(np.float32)         # an elevation raster (DEM_ELEV modality)
    p = str(tmp_path / "f32.tif")
    Image.fromarray(dem[0]).save(p, format="TIFF", compression="tiff_lzw")
    got, info = rio.read_raster(p)
    assert info.dtype == np.float32 and np.array_equal(got, dem)
    # GDAL's PREDICTOR=3 for elevation rasters (floating-point predictor: byte planes + byte differences), libtiff-written
    smooth = (np.cumsum(rng.standard_normal((1, 211, 333)), axis=2) * 3 + 400).astype(np.float32)
    for comp in ("tiff_lzw", "tiff_adobe_deflate"):
        p = str(tmp_path / "fp3.tif")
        Image.fromarray(smooth[0]).save(p, format="TIFF", compression=comp, tiffinfo={317: 3})
        got, info = rio.read_raster(p)
        assert info.predictor == 3 and info.dtype == np.float32 and np.array_equal(got, smooth)
        assert np.array_equal(rio.read_window(p, 100, -5, 50, 60)[0, :, 5:], smooth[0, 100:150, :55])
    # and written here (tiled float32 / uint16), read by libtiff
    for arr in (dem, a16):
        p = str(tmp_path / "w.tif")
        rio.write_geotiff(p, arr, LEFT, TOP, 1.0, epsg=2154, compression="deflate", block=128)
        assert np.array_equal(pillow_frames(p)[0], arr[0])
        got, info = rio.read_raster(p)
        assert info.dtype == arr.dtype and np.array_equal(got, arr)


def test_pixel_interleaved_and_band_interleaved_multiband(tmp_path):
    rng = np.random.default_rng(7)
    rgbi = class_map(rng, 530, 700, 4, n_cls=256)
    p = str(tmp_path / "rgbi.tif")
    rio.write_geotiff(p, rgbi, LEFT, TOP, RES, epsg=2154, pixel_interleave=True, predictor=2)     # GDAL's usual ortho layout
    # libtiff (tiled, chunky, predictor): Pillow maps RGB + one unspecified extra sample to RGB and drops the 4th band, so the
    # infrared band is checked through a second file with the bands rolled
    assert np.array_equal(pillow_frames(p)[0].transpose(2, 0, 1), rgbi[:3])
    rolled = str(tmp_path / "irgb.tif")
    rio.write_geotiff(rolled, rgbi[[3, 0, 1, 2]], pixel_interleave=True, predictor=2)
    assert np.array_equal(pillow_frames(rolled)[0].transpose(2, 0, 1), rgbi[[3, 0, 1]])
    got, info = rio.read_raster(p)
    assert info.planar == 1 and info.count == 4 and np.array_equal(got, rgbi)
    prob = rng.integers(0, 256, (19, 300, 260)).astype(np.uint8)                                  # class_prob: 19 bands
    for chunky in (False, True):
        p = str(tmp_path / f"prob{int(chunky)}.tif")
        rio.write_geotiff(p, prob, LEFT, TOP, RES, compression="deflate", pixel_interleave=chunky, block=128)
        got, info = rio.read_raster(p)
        assert info.planar == (1 if chunky else 2) and info.count == 19 and np.array_equal(got, prob)
        sel = rio.read_window(p, 100, 50, 64, 64, bands=[19, 3, 3])
        assert np.array_equal(sel, prob[[18, 2, 2], 100:164, 50:114])
    # non-contiguous input: a band-subset view and a cropped view are written without a copy being needed by the caller
    view = prob[::2, 10:250, 5:200]
    p = str(tmp_path / "view.tif")
    rio.write_geotiff(p, view, compression="lzw", block=64)
    assert np.array_equal(rio.read_raster(p)[0], view)


def test_boundless_windows_equal_zero_padded_slicing(tmp_path):
    """dataset.py:108-115: windows hanging over any edge, far outside, empty; tiled and striped files."""
    rng = np.random.default_rng(8)
    a = rng.integers(1, 256, (4, 700, 900)).astype(np.uint8)
    tiled, strips = str(tmp_path / "t.tif"), str(tmp_path / "s.tif")
    rio.write_geotiff(tiled, a, LEFT, TOP, RES, pixel_interleave=True, block=128)
    Image.fromarray(a.transpose(1, 2, 0), "RGBA").save(strips, format="TIFF", compression="tiff_lzw")
    pad = 600
    padded = np.zeros((4, 700 + 2 * pad, 900 + 2 * pad), np.uint8)
    padded[:, pad:pad + 700, pad:pad + 900] = a
    windows = [(-64, -64, 512, 512), (300, 500, 512, 512), (188, 388, 512, 512), (-600, -600, 100, 100), (699, 899, 1, 1),
               (700, 0, 10, 10), (0, 0, 700, 900), (-10, -10, 720, 920), (5, 5, 0, 10)]
    windows += [(int(rng.integers(-500, 700)), int(rng.integers(-500, 900)), int(rng.integers(1, 600)), int(rng.integers(1, 600)))
                for _ in range(25)]
    for path in (tiled, strips):
        for r0, c0, h, w in windows:
            got = rio.read_window(path, r0, c0, h, w)
            assert np.array_equal(got, padded[:, pad + r0:pad + r0 + h, pad + c0:pad + c0 + w]), (path, r0, c0, h, w)
    # decoding straight into a larger row-strided buffer (a slab of a pinned raster)
    big = np.full((4, 600, 1000), 9, np.uint8)
    rio.read_window(tiled, 100, 200, 300, 400, out=big[:, 50:350, 100:500])
    assert np.array_equal(big[:, 50:350, 100:500], a[:, 100:400, 200:600]) and big[0, 0, 0] == 9 and big[3, 350, 100] == 9
    with pytest.raises(rio.RasterIOError, match="band index"):
        rio.read_window(tiled, 0, 0, 8, 8, bands=[5])


def test_bigtiff_and_overviews_and_cog_layout(tmp_path):
    rng = np.random.default_rng(9)
    a = class_map(rng, 1500, 2100)
    p = str(tmp_path / "big.tif")
    rio.write_geotiff(p, a, LEFT, TOP, RES, epsg=2154, bigtiff=1, overviews=2, block=256)
    info = rio.tiff_info(p)
    assert info.bigtiff and info.overviews == 2 and open(p, "rb").read(4) == b"II+\x00"
    assert np.array_equal(rio.read_raster(p)[0], a)
    # GDAL's nearest rule (postprocess.py:44): source pixel floor(0.5 + i * src / dst), level by level
    def nearest(x):
        h, w = x.shape[-2:]
        dh, dw = (h + 1) // 2, (w + 1) // 2
        ys = np.minimum((0.5 + np.arange(dh) * (h / dh)).astype(np.int64), h - 1)
        xs = np.minimum((0.5 + np.arange(dw) * (w / dw)).astype(np.int64), w - 1)
        return x[..., ys[:, None], xs[None, :]]
    o1, o2 = nearest(a), nearest(nearest(a))
    assert np.array_equal(rio.read_raster(p, level=1)[0], o1) and np.array_equal(rio.read_raster(p, level=2)[0], o2)
    assert rio.tiff_info(p, 2).width == 525 and not rio.tiff_info(p, 1).has_georef
    with pytest.raises(rio.RasterIOError, match="no overview level 3"):
        rio.tiff_info(p, 3)

    # COG (classic TIFF): libtiff sees 1 + n images, IFDs sit in front of all pixel data, smallest overview first
    cog = str(tmp_path / "cog.tif")
    rio.write_geotiff(cog, a, LEFT, TOP, RES, epsg=2154, cog=True, overviews=-1)
    frames = pillow_frames(cog)
    assert [f.shape for f in frames] == [(1500, 2100), (750, 1050), (375, 525), (188, 263)]   # halve until <= 512 on both sides
    assert np.array_equal(frames[0], a[0]) and np.array_equal(frames[1], o1[0]) and np.array_equal(frames[2], o2[0])
    assert np.array_equal(frames[3], nearest(o2)[0])
    raw = open(cog, "rb").read()
    assert raw[8:8 + 43] == b"GDAL_STRUCTURAL_METADATA_SIZE=000140 bytes\n" and b"LAYOUT=IFDS_BEFORE_DATA" in raw[:200]
    offs = []
    with Image.open(cog) as im:
        for k in range(4):
            im.seek(k)
            offs.append((min(im.tag_v2[324]), max(im.tag_v2[324]), list(im.tag_v2[324]), list(im.tag_v2[325])))
    assert offs[3][1] < offs[2][0] and offs[2][1] < offs[1][0] and offs[1][1] < offs[0][0]                           # overview data before full resolution
    for _, _, o, c in offs:                                                              # size leader / repeated-bytes trailer
        for off, cnt in zip(o, c):
            assert struct.unpack("<I", raw[off - 4:off])[0] == cnt and raw[off + cnt:off + cnt + 4] == raw[off + cnt - 4:off + cnt]
    mode = str(tmp_path / "mode.tif")
    rio.write_geotiff(mode, a, overviews=1, overview_resampling="mode")
    m = rio.read_raster(mode, level=1)[0][0]
    blocks = a[0].reshape(750, 2, 1050, 2).transpose(0, 2, 1, 3).reshape(750, 1050, 4)
    want = np.array([[np.bincount(q, minlength=19).argmax() for q in row] for row in blocks[:40]])   # ties -> smallest value
    assert np.array_equal(m[:40], want)


def test_convert_to_cog_like_the_reference(tmp_path):
    """postprocess.py:33-52 / inference.py:633-641 through the product mirror: LZW, 512 blocks, nearest overviews, the input
    file removed, FileNotFoundError for a missing input."""
    from flair_for_aigle_b200.flair_zonal_detection.inference import postpro_outputs
    from flair_for_aigle_b200.flair_zonal_detection.postprocess import convert_to_cog
    rng = np.random.default_rng(10)
    a = class_map(rng, 1300, 1100)
    src = str(tmp_path / "zone_task_argmax_i.tif")
    Image.fromarray(a[0]).save(src, format="TIFF", compression="tiff_lzw")               # any TIFF in ...
    with pytest.raises(FileNotFoundError, match="Input file not found"):
        convert_to_cog(str(tmp_path / "nope.tif"), str(tmp_path / "out.tif"))
    from flair_for_aigle_b200.flair_zonal_detection.geotiff import write_geotiff
    write_geotiff(src, a, LEFT, TOP, RES, "EPSG:2154")
    postpro_outputs({"task": src}, {"cog_conversion": False})
    assert os.path.isfile(src)
    postpro_outputs({"task": src}, {"cog_conversion": True})
    cog = src.replace(".tif", "_COG.tif")
    assert os.path.isfile(cog) and not os.path.exists(src)                               # ... COG out, input removed
    info = rio.tiff_info(cog)
    assert info.tiled and info.block_w == info.block_h == 512 and info.compression == rio.COMP_LZW and info.overviews == 2
    assert (info.left, info.top, info.res_x, info.epsg) == (LEFT, TOP, RES, 2154)
    frames = pillow_frames(cog)
    assert np.array_equal(frames[0], a[0]) and [f.shape for f in frames] == [(1300, 1100), (650, 550), (325, 275)]


def _handmade_big_endian_tiff(path, arr16):
    """An uncompressed big-endian ('MM') uint16 strip file with a PixelIsPoint GeoKey, built byte by byte."""
    h, w = arr16.shape
    data = arr16.astype(">u2").tobytes()
    scale = struct.pack(">3d", 2.0, 2.0, 0.0)
    tie = struct.pack(">6d", 0.0, 0.0, 0.0, 1000.0, 5000.0, 0.0)
    keys = struct.pack(">16H", 1, 1, 0, 3, 1024, 0, 1, 2, 1025, 0, 1, 2, 2048, 0, 1, 4326)
    pos = 8 + len(data)
    blobs = {}
    for tag, b in ((33550, scale), (33922, tie), (34735, keys)):
        blobs[tag] = pos
        pos += len(b)
    ents = [(256, 3, 1, struct.pack(">HH", w, 0)), (257, 3, 1, struct.pack(">HH", h, 0)), (258, 3, 1, struct.pack(">HH", 16, 0)),
            (259, 3, 1, struct.pack(">HH", 1, 0)), (262, 3, 1, struct.pack(">HH", 1, 0)), (273, 4, 1, struct.pack(">I", 8)),
            (277, 3, 1, struct.pack(">HH", 1, 0)), (278, 3, 1, struct.pack(">HH", h, 0)), (279, 4, 1, struct.pack(">I", len(data))),
            (33550, 12, 3, struct.pack(">I", blobs[33550])), (33922, 12, 6, struct.pack(">I", blobs[33922])),
            (34735, 3, 16, struct.pack(">I", blobs[34735]))]
    with open(path, "wb") as f:
        f.write(b"MM" + struct.pack(">HI", 42, pos) + data + scale + tie + keys)
        f.write(struct.pack(">H", len(ents)) + b"".join(struct.pack(">HHI", t, ty, n) + v for t, ty, n, v in ents) + struct.pack(">I", 0))


def test_big_endian_file_and_pixel_is_point(tmp_path):
    rng = np.random.default_rng(11)
    a = rng.integers(0, 65536, (37, 53)).astype(np.uint16)
    p = str(tmp_path / "mm.tif")
    _handmade_big_endian_tiff(p, a)
    with Image.open(p) as im:
        assert np.array_equal(np.asarray(im), a)                         # libtiff agrees the file is well formed
    got, info = rio.read_raster(p)
    assert np.array_equal(got[0], a) and info.dtype == np.uint16
    # PixelIsPoint: the tie point is the CENTRE of pixel (0, 0) -> the outer corner lies half a pixel up-left (GDAL's rule)
    assert (info.left, info.top, info.res_x) == (999.0, 5001.0, 2.0) and info.geographic and info.epsg == 4326


def test_errors_are_loud(tmp_path):
    with pytest.raises(rio.RasterIOError, match="cannot open"):
        rio.tiff_info(str(tmp_path / "missing.tif"))
    bad = tmp_path / "bad.tif"
    bad.write_bytes(b"not a tiff at all")
    with pytest.raises(rio.RasterIOError, match="not a TIFF"):
        rio.tiff_info(str(bad))
    jpeg = str(tmp_path / "jpeg.tif")
    Image.fromarray(np.zeros((64, 64, 3), np.uint8)).save(jpeg, format="TIFF", compression="jpeg")
    assert rio.tiff_info(jpeg).compression == 7
    with pytest.raises(rio.RasterIOError, match="compression 7 is not supported"):
        rio.read_raster(jpeg)
    with pytest.raises(rio.RasterIOError, match="multiples of 16"):
        rio.write_geotiff(str(tmp_path / "x.tif"), np.zeros((1, 8, 8), np.uint8), block=100)
    with pytest.raises(rio.RasterIOError, match="float64"):
        rio.write_geotiff(str(tmp_path / "x.tif"), np.zeros((1, 8, 8), np.float64))
    trunc = str(tmp_path / "trunc.tif")
    rio.write_geotiff(trunc, np.ones((1, 600, 600), np.uint8) * 3, compression="none")
    raw = open(trunc, "rb").read()
    open(trunc, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(rio.RasterIOError, match="outside the file"):
        rio.read_raster(trunc)


def test_corrupted_files_never_crash_the_process(tmp_path):
    """A file parser meets broken files: 400 random corruptions of valid TIFFs (bytes flipped in the header, the directory,
    the block tables and the compressed data; truncations) either decode to SOMETHING of the right shape or raise
    RasterIOError -- no crash, no hang, no out-of-bounds write (the destination carries guard rows)."""
    rng = np.random.default_rng(12)
    a = class_map(rng, 300, 260, 3)
    seeds = []
    for k, kw in enumerate((dict(compression="lzw", predictor=2, pixel_interleave=True, block=64),
                            dict(compression="deflate", block=128, overviews=1), dict(compression="none", block=64, bigtiff=1))):
        p = str(tmp_path / f"seed{k}.tif")
        rio.write_geotiff(p, a, LEFT, TOP, RES, epsg=2154, **kw)
        seeds.append(open(p, "rb").read())
    outcomes = {"ok": 0, "error": 0}
    victim = str(tmp_path / "victim.tif")
    for trial in range(400):
        raw = bytearray(seeds[trial % 3])
        kind = trial % 4
        if kind == 0:                                         # header + first directory region
            for _ in range(int(rng.integers(1, 6))):
                raw[int(rng.integers(0, min(400, len(raw))))] = int(rng.integers(0, 256))
        elif kind == 1:                                       # anywhere
            for _ in range(int(rng.integers(1, 40))):
                raw[int(rng.integers(0, len(raw)))] = int(rng.integers(0, 256))
        elif kind == 2:                                       # truncation
            raw = raw[:int(rng.integers(8, len(raw)))]
        else:                                                 # a run of garbage
            at = int(rng.integers(0, len(raw) - 64))
            raw[at:at + 64] = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        with open(victim, "wb") as f:
            f.write(bytes(raw))
        try:
            info = rio.tiff_info(victim)
            if info.width * info.height * info.count * info.dtype.itemsize > 64 << 20:
                raise rio.RasterIOError("implausibly large")  # a corrupted size field: the caller would refuse it as well
            guard = np.full((info.count, info.height + 2, info.width), 0xAB, info.dtype.str.replace("f4", "u4"))
            out = guard[:, 1:-1].view(info.dtype)
            rio.read_window(victim, 0, 0, info.height, info.width, out=out, info=info)
            assert (guard[:, 0].view(np.uint8) == 0xAB).all() and (guard[:, -1].view(np.uint8) == 0xAB).all()
            outcomes["ok"] += 1
        except rio.RasterIOError:
            outcomes["error"] += 1
    assert outcomes["ok"] + outcomes["error"] == 400 and outcomes["error"] > 50 and outcomes["ok"] > 20, outcomes


def test_random_layouts_round_trip_and_decode_in_libtiff(tmp_path):
    """Property test (hypothesis): any shape x band count x dtype x block size x codec x predictor x interleave x BigTIFF x
    overview count round-trips bit for bit through the writer and the reader, windows included, and libtiff agrees on the
    single-band / RGB cases it can open."""
    from hypothesis import HealthCheck, given, settings, strategies as st

    @settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
    @given(h=st.integers(1, 300), w=st.integers(1, 300), count=st.sampled_from([1, 1, 2, 3, 4, 7]),
           dtype=st.sampled_from(["u1", "u1", "u1", "u2", "f4", "i2"]), block=st.sampled_from([16, 32, 64, 128]),
           comp=st.sampled_from(["lzw", "deflate", "none"]), pred2=st.booleans(), chunky=st.booleans(), big=st.booleans(),
           ovr=st.integers(0, 2), seed=st.integers(0, 10 ** 6), smooth=st.booleans())
    def check(h, w, count, dtype, block, comp, pred2, chunky, big, ovr, seed, smooth):
        rng = np.random.default_rng(seed)
        a = rng.integers(0, 200, (count, h, w)).astype(dtype)
        if smooth:
            a = np.sort(a, axis=2)
        predictor = 2 if (pred2 and dtype == "u1" and comp != "none") else 1
        p = str(tmp_path / "h.tif")
        rio.write_geotiff(p, a, LEFT, TOP, RES, epsg=2154, compression=comp, predictor=predictor, block=block,
                          pixel_interleave=chunky, bigtiff=int(big), overviews=ovr)
        got, info = rio.read_raster(p)
        assert np.array_equal(got, a) and info.dtype == a.dtype and info.bigtiff == big and info.count == count
        assert info.overviews == min(ovr, sum(1 for k in range(ovr) if max((h + (1 << k) - 1) >> k, (w + (1 << k) - 1) >> k) > 1))
        r0, c0 = int(rng.integers(-20, h)), int(rng.integers(-20, w))
        hh, ww = int(rng.integers(1, h + 30)), int(rng.integers(1, w + 30))
        pad = np.zeros((count, h + 800, w + 800), a.dtype)
        pad[:, 400:400 + h, 400:400 + w] = a
        assert np.array_equal(rio.read_window(p, r0, c0, hh, ww), pad[:, 400 + r0:400 + r0 + hh, 400 + c0:400 + c0 + ww])
        if not big and dtype == "u1" and (count == 1 or (count == 3 and chunky)):
            with Image.open(p) as im:
                lib = np.asarray(im)
            assert np.array_equal(lib if count == 1 else lib.transpose(2, 0, 1), a[0] if count == 1 else a)
    check()


def test_host_thread_share_in_a_multi_gpu_job(monkeypatch):
    """One process per GPU: each rank's file codecs take their share of the cores, not all of them."""
    import os
    monkeypatch.delenv("FZ_IO_THREADS", raising=False)
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    assert rio.host_threads(0) == 0 and rio.host_threads(3) == 3
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    monkeypatch.setattr(os, "cpu_count", lambda: 96)
    assert rio.host_threads(0) == 12 and rio.host_threads(5) == 5
    monkeypatch.setattr(os, "cpu_count", lambda: 4)
    assert rio.host_threads(0) == 1
    monkeypatch.setenv("FZ_IO_THREADS", "6")
    assert rio.host_threads(0) == 6
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    monkeypatch.delenv("FZ_IO_THREADS")
    assert rio.host_threads(0) == 0


def test_host_benchmark_tool_runs(tmp_path, monkeypatch):
    """tools/raster_io_bench.py (the source of profiles/r2_raster_io_bench.txt) at a toy size: every leg runs and verifies
    its own round trips."""
    import runpy
    import sys
    out = str(tmp_path / "bench.txt")
    monkeypatch.setattr(sys, "argv", ["raster_io_bench.py", "--size", "700", "--reps", "1", "--out", out])
    runpy.run_path(os.path.join(ROOT, "tools", "raster_io_bench.py"), run_name="__main__")
    text = open(out).read()
    assert "WRITE argmax raster" in text and "READ 4-band uint8 ortho" in text and "READ JPEG 2000 ortho" in text
    assert "skipped" not in text

"""GeoTIFF I/O at both ends of the path (flair_zonal_detection/geotiff.py) -- host code, no GPU."""
import os

import numpy as np
import pytest

from flair_for_aigle_b200.flair_zonal_detection.geotiff import read_geotiff, write_geotiff
from flair_for_aigle_b200.flair_zonal_detection.raster import open_raster

L, T, RES = 700000.0, 6600000.0, 0.2


def test_single_band_lzw_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    arr = rng.integers(0, 19, (1, 777, 1301), dtype=np.uint8)
    p = str(tmp_path / "classes.tif")
    write_geotiff(p, arr, L, T, RES, "EPSG:2154")
    got, left, top, res, crs = read_geotiff(p)
    assert np.array_equal(got, arr) and (left, top, res, crs) == (L, T, RES, "EPSG:2154")
    from PIL import Image
    with Image.open(p) as im:                      # the reference's profile: compress='lzw'
        assert im.info.get("compression") == "tiff_lzw"
        assert tuple(im.tag_v2[33550])[:2] == (RES, RES) and tuple(im.tag_v2[33922])[3:5] == (L, T)


def test_multi_band_planar_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    arr = rng.integers(0, 256, (19, 300, 517), dtype=np.uint8)
    p = str(tmp_path / "probs.tif")
    write_geotiff(p, arr, L, T, RES, "EPSG:2154")
    got, left, top, res, crs = read_geotiff(p)
    assert np.array_equal(got, arr) and (left, top, res, crs) == (L, T, RES, "EPSG:2154")


def test_open_raster_reads_a_4_band_geotiff(tmp_path):
    """An RGBI uint8 GeoTIFF as a GIS tool writes it (pixel-interleaved, LZW) -> ZoneRaster with the right georeference."""
    from PIL import Image, TiffImagePlugin
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (640, 900, 4), dtype=np.uint8)
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[33550] = (RES, RES, 0.0)
    ifd.tagtype[33550] = 12
    ifd[33922] = (0.0, 0.0, 0.0, L, T, 0.0)
    ifd.tagtype[33922] = 12
    p = str(tmp_path / "ortho.tif")
    Image.fromarray(img, mode="RGBA").save(p, format="TIFF", compression="tiff_lzw", tiffinfo=ifd)
    r = open_raster(p)
    assert (r.count, r.height, r.width) == (4, 640, 900)
    assert np.array_equal(r.read(), img.transpose(2, 0, 1))
    b = r.bounds
    assert (b.left, b.top) == (L, T) and abs(b.right - (L + 900 * RES)) < 1e-6 and abs(b.bottom - (T - 640 * RES)) < 1e-6


def test_rejects_tiff_without_georeference(tmp_path):
    from PIL import Image
    p = str(tmp_path / "plain.tif")
    Image.fromarray(np.zeros((8, 8), np.uint8)).save(p)
    with pytest.raises(ValueError):
        read_geotiff(p)
    # ... unless a world file (pixel-centre convention) or a MapInfo .tab (pixel corners) sits next to it
    open(str(tmp_path / "plain.tfw"), "w").write(f"{RES}\n0\n0\n-{RES}\n{L + RES / 2!r}\n{T - RES / 2!r}\n")
    got, left, top, res, crs = read_geotiff(p)
    assert got.shape == (1, 8, 8) and abs(left - L) < 1e-6 and abs(top - T) < 1e-6 and res == RES and crs is None
    r = open_raster(p)
    assert abs(r.bounds.left - L) < 1e-6 and abs(r.bounds.bottom - (T - 8 * RES)) < 1e-6
    os.remove(str(tmp_path / "plain.tfw"))
    open(str(tmp_path / "plain.tab"), "w").write(
        f'Definition Table\n  File "plain.tif"\n  Type "RASTER"\n  ({L!r},{T!r}) (0,0) Label "Pt 1",\n'
        f'  ({L + 8 * RES!r},{T!r}) (8,0) Label "Pt 2",\n  ({L!r},{T - 8 * RES!r}) (0,8) Label "Pt 3"\n'
        '  CoordSys Earth Projection 3, 33, "m", 3, 46.5, 44, 49, 700000, 6600000\n')
    _, left, top, res, crs = read_geotiff(p)
    assert abs(left - L) < 1e-6 and abs(top - T) < 1e-6 and abs(res - RES) < 1e-9 and crs == "EPSG:2154"


def test_crs_codes_are_parsed_not_guessed():
    """Round-1 advisor finding: the EPSG code used to be 'the first 4-6 digit run' of the CRS string, which reads 1980 out of
    'GRS 1980' in a WKT, and geographic CRSs were written as projected ones."""
    from flair_for_aigle_b200.flair_zonal_detection.geotiff import _epsg, _geokeys
    wkt_l93 = ('PROJCS["RGF93 / Lambert-93",GEOGCS["RGF93",DATUM["Reseau_Geodesique_Francais_1993",SPHEROID["GRS 1980",6378137,'
               '298.257222101,AUTHORITY["EPSG","7019"]],AUTHORITY["EPSG","6171"]],AUTHORITY["EPSG","4171"]],'
               'PROJECTION["Lambert_Conformal_Conic_2SP"],UNIT["metre",1],AUTHORITY["EPSG","2154"]]')
    wkt2_wgs = 'GEOGCRS["WGS 84",DATUM["World Geodetic System 1984",ELLIPSOID["WGS 84",6378137,298.257223563]],ID["EPSG",4326]]'
    assert _epsg("EPSG:2154") == 2154 and _epsg("epsg:4326") == 4326 and _epsg("2154") == 2154
    assert _epsg("urn:ogc:def:crs:EPSG::2154") == 2154
    assert _epsg(wkt_l93) == 2154 and _epsg(wkt2_wgs) == 4326
    assert _epsg('PROJCS["custom",GEOGCS["x",DATUM["d",SPHEROID["GRS 1980",6378137,298.25]]]]') is None     # no authority
    assert _epsg(None) is None and _epsg("+proj=lcc +lat_1=49 +ellps=GRS80") is None

    def keys(crs):
        flat = _geokeys(crs)
        return {flat[i]: flat[i + 3] for i in range(4, len(flat), 4)}
    assert keys("EPSG:2154") == {1024: 1, 1025: 1, 3072: 2154}          # projected: ProjectedCSTypeGeoKey
    assert keys("EPSG:4326") == {1024: 2, 1025: 1, 2048: 4326}          # geographic: GeographicTypeGeoKey
    assert keys(wkt2_wgs) == {1024: 2, 1025: 1, 2048: 4326} and keys(wkt_l93) == {1024: 1, 1025: 1, 3072: 2154}
    assert keys(None) == {1024: 1, 1025: 1}


def test_geographic_crs_roundtrip(tmp_path):
    from flair_for_aigle_b200.flair_zonal_detection.geotiff import read_geotiff, write_geotiff
    arr = (np.arange(6 * 7, dtype=np.uint8).reshape(1, 6, 7) * 3)
    p = write_geotiff(str(tmp_path / "geo.tif"), arr, 2.25, 48.75, 0.0001, "EPSG:4326")
    got, left, top, res, crs = read_geotiff(p)
    assert np.array_equal(got, arr) and (left, top, res, crs) == (2.25, 48.75, 0.0001, "EPSG:4326")


def test_outputs_are_tiled_lzw_and_the_geokeys_are_the_documented_ones(tmp_path):
    """The writer is libfz_rasterio.so: 512 x 512 LZW tiles for the argmax raster AND the class_prob planes (the reference's
    compress='lzw' profile for both, inference.py:182-203); the GeoKeyDirectory in the file is `_geokeys(crs)`."""
    from PIL import Image
    from flair_for_aigle_b200 import raster_io as rio
    from flair_for_aigle_b200.flair_zonal_detection.geotiff import _geokeys
    rng = np.random.default_rng(5)
    for count, crs in ((1, "EPSG:2154"), (19, "EPSG:4326"), (1, None)):
        arr = rng.integers(0, 19, (count, 600, 700), dtype=np.uint8)
        p = write_geotiff(str(tmp_path / f"o{count}.tif"), arr, L, T, RES, crs)
        info = rio.tiff_info(p)
        assert info.tiled and (info.block_w, info.block_h) == (512, 512) and info.compression == rio.COMP_LZW
        assert info.planar == (2 if count > 1 else 1) and info.count == count and not info.bigtiff
        with Image.open(p) as im:
            assert tuple(im.tag_v2[34735]) == _geokeys(crs)
        got, left, top, res, crs_back = read_geotiff(p)
        assert np.array_equal(got, arr) and (left, top, res, crs_back) == (L, T, RES, crs)


def test_float32_elevation_geotiff_and_jpeg_in_tiff_fallback(tmp_path):
    from PIL import Image, TiffImagePlugin
    from flair_for_aigle_b200 import raster_io as rio
    rng = np.random.default_rng(6)
    dem = (rng.standard_normal((1, 300, 400)) * 30 + 250).astype(np.float32)
    p = str(tmp_path / "dem.tif")
    rio.write_geotiff(p, dem, L, T, 1.0, epsg=2154, compression="deflate")
    r = open_raster(p)
    assert r.read().dtype == np.float32 and np.array_equal(r.read(), dem) and r.res == (1.0, 1.0) and r.crs == "EPSG:2154"
    # a codec the library does not implement (JPEG-in-TIFF): libtiff through Pillow decodes it
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[33550], ifd.tagtype[33550] = (RES, RES, 0.0), 12
    ifd[33922], ifd.tagtype[33922] = (0.0, 0.0, 0.0, L, T, 0.0), 12
    img = np.kron(rng.integers(0, 256, (20, 30, 3)).astype(np.uint8), np.ones((16, 16, 1), np.uint8))
    p = str(tmp_path / "jpeg.tif")
    Image.fromarray(img).save(p, format="TIFF", compression="jpeg", tiffinfo=ifd)
    got, left, top, res, _ = read_geotiff(p)
    with Image.open(p) as im:
        want = np.asarray(im).transpose(2, 0, 1)
    assert np.array_equal(got, want) and (left, top, res) == (L, T, RES)
    assert np.abs(got.astype(int) - img.transpose(2, 0, 1)).mean() < 6          # lossy, but the right picture


def _jp2_with_box(path, arr_hwc, box: bytes):
    """Lossless JPEG 2000 written by OpenJPEG (Pillow), with one extra box spliced in front of the codestream."""
    import struct
    from PIL import Image
    Image.fromarray(arr_hwc).save(path, format="JPEG2000", irreversible=False)
    raw = open(path, "rb").read()
    at = raw.index(b"jp2c") - 4
    open(path, "wb").write(raw[:at] + box + raw[at:])


def _geojp2_box(left, top, res, epsg, point=False):
    import struct
    scale = struct.pack("<3d", res, res, 0.0)
    tie = struct.pack("<6d", 0.0, 0.0, 0.0, left + (0.5 * res if point else 0.0), top - (0.5 * res if point else 0.0), 0.0)
    keys = struct.pack("<16H", 1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 2 if point else 1, 3072, 0, 1, epsg)
    # degenerate 1 x 1 GeoTIFF, little endian: header, one pixel, out-of-line values, directory
    body = b"\x00"
    pos = 8 + len(body) + 1
    offs = {}
    blob = b""
    for tag, b in ((33550, scale), (33922, tie), (34735, keys)):
        offs[tag] = pos + len(blob)
        blob += b
    ents = [(256, 3, 1, struct.pack("<HH", 1, 0)), (257, 3, 1, struct.pack("<HH", 1, 0)), (258, 3, 1, struct.pack("<HH", 8, 0)),
            (259, 3, 1, struct.pack("<HH", 1, 0)), (262, 3, 1, struct.pack("<HH", 1, 0)), (273, 4, 1, struct.pack("<I", 8)),
            (277, 3, 1, struct.pack("<HH", 1, 0)), (278, 3, 1, struct.pack("<HH", 1, 0)), (279, 4, 1, struct.pack("<I", 1)),
            (33550, 12, 3, struct.pack("<I", offs[33550])), (33922, 12, 6, struct.pack("<I", offs[33922])),
            (34735, 3, 16, struct.pack("<I", offs[34735]))]
    ifd_at = pos + len(blob)
    tiff = b"II*\x00" + struct.pack("<I", ifd_at) + body + b"\x00" + blob
    tiff += struct.pack("<H", len(ents)) + b"".join(struct.pack("<HHI", t, ty, n) + v for t, ty, n, v in ents) + struct.pack("<I", 0)
    payload = bytes.fromhex("b14bf8bd083d4b43a5ae8cd7d5a6ce03") + tiff
    return struct.pack(">I4s", 8 + len(payload), b"uuid") + payload


def test_jpeg2000_inputs_with_geojp2_gmljp2_and_world_file(tmp_path):
    """The reference's product script feeds *.jp2 orthos (inference.py:60, scripts/run_fast_aigle_segmentation.py:88)."""
    import struct
    from PIL import features
    from flair_for_aigle_b200.flair_zonal_detection.geotiff import read_jp2
    if not features.check("jpg_2000"):
        pytest.skip("Pillow without OpenJPEG")
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (300, 420, 4), dtype=np.uint8)                 # RGBI
    want = img.transpose(2, 0, 1)
    p = str(tmp_path / "geojp2.jp2")
    _jp2_with_box(p, img, _geojp2_box(L, T, RES, 2154))
    got, left, top, res, crs = read_jp2(p)
    assert np.array_equal(got, want) and (left, top, res, crs) == (L, T, RES, "EPSG:2154")
    r = open_raster(p)                                                        # ... and through the rasterio.open stand-in
    assert (r.count, r.height, r.width) == (4, 300, 420) and np.array_equal(r.read(), want) and r.bounds.left == L
    p = str(tmp_path / "point.jp2")
    _jp2_with_box(p, img, _geojp2_box(L, T, RES, 2154, point=True))           # PixelIsPoint tie point = pixel centre
    _, left, top, res, _ = read_jp2(p)
    assert abs(left - L) < 1e-9 and abs(top - T) < 1e-9
    # GMLJP2: asoc(lbl, asoc(lbl, xml)) with a RectifiedGrid whose origin is the centre of the first pixel
    gml = (f'<gml:FeatureCollection xmlns:gml="http://www.opengis.net/gml"><gml:featureMember><gml:RectifiedGridCoverage>'
           f'<gml:rectifiedGridDomain><gml:RectifiedGrid dimension="2" srsName="urn:ogc:def:crs:EPSG::2154"><gml:limits><gml:GridEnvelope>'
           f'<gml:low>0 0</gml:low><gml:high>419 299</gml:high></gml:GridEnvelope></gml:limits>'
           f'<gml:origin><gml:Point srsName="urn:ogc:def:crs:EPSG::2154"><gml:pos>{L + RES / 2!r} {T - RES / 2!r}</gml:pos></gml:Point></gml:origin>'
           f'<gml:offsetVector srsName="urn:ogc:def:crs:EPSG::2154">{RES} 0</gml:offsetVector>'
           f'<gml:offsetVector srsName="urn:ogc:def:crs:EPSG::2154">0 -{RES}</gml:offsetVector>'
           f'</gml:RectifiedGrid></gml:rectifiedGridDomain></gml:RectifiedGridCoverage></gml:featureMember></gml:FeatureCollection>').encode()

    def box(kind, payload):
        return struct.pack(">I4s", 8 + len(payload), kind) + payload
    asoc = box(b"asoc", box(b"lbl ", b"gml.data") + box(b"asoc", box(b"lbl ", b"gml.root-instance") + box(b"xml ", gml)))
    p = str(tmp_path / "gml.jp2")
    _jp2_with_box(p, img[:, :, :3], asoc)
    got, left, top, res, crs = read_jp2(p)
    assert np.array_equal(got, want[:3]) and abs(left - L) < 1e-6 and abs(top - T) < 1e-6 and res == RES and crs == "EPSG:2154"
    # world file
    p = str(tmp_path / "wf.jp2")
    _jp2_with_box(p, img[:, :, 0], b"")
    with pytest.raises(ValueError, match="no georeferencing"):
        read_jp2(p)
    with pytest.raises(ValueError):                                           # open_raster: nothing else can read it either
        open_raster(p)
    open(str(tmp_path / "wf.j2w"), "w").write(f"{RES}\n0.0\n0.0\n-{RES}\n{L + RES / 2!r}\n{T - RES / 2!r}\n")
    got, left, top, res, crs = read_jp2(p)
    assert np.array_equal(got[0], img[:, :, 0]) and abs(left - L) < 1e-6 and abs(top - T) < 1e-6 and crs is None
    bad = tmp_path / "bad.jp2"
    bad.write_bytes(b"\x00" * 64)
    with pytest.raises(ValueError, match="not a JP2"):
        read_jp2(str(bad))
    # a bare codestream (.j2k) has no boxes: side files only
    from PIL import Image
    p = str(tmp_path / "raw.j2k")
    Image.fromarray(img[:, :, 1]).save(p, format="JPEG2000", irreversible=False)
    with pytest.raises(ValueError, match="codestream carries no georeferencing"):
        read_jp2(p)
    open(str(tmp_path / "raw.j2w"), "w").write(f"{RES}\n0.0\n0.0\n-{RES}\n{L + RES / 2!r}\n{T - RES / 2!r}\n")
    got, left, top, res, crs = read_jp2(p)
    assert np.array_equal(got[0], img[:, :, 1]) and abs(left - L) < 1e-6 and abs(top - T) < 1e-6
    assert open_raster(p).shape == (300, 420)
    # MapInfo .tab registration (IGN's BD ORTHO deliveries): corner control points + the Lambert-93 CoordSys clause
    p = str(tmp_path / "tab.jp2")
    _jp2_with_box(p, img[:, :, 0], b"")
    right, bottom = L + 420 * RES, T - 300 * RES
    open(str(tmp_path / "tab.tab"), "w", encoding="latin-1").write(
        '!table\n!version 300\n!charset WindowsLatin1\n\nDefinition Table\n  File "tab.jp2"\n  Type "RASTER"\n'
        f'  ({L!r},{T!r}) (0,0) Label "Pt 1",\n  ({right!r},{T!r}) (420,0) Label "Pt 2",\n'
        f'  ({right!r},{bottom!r}) (420,300) Label "Pt 3",\n  ({L!r},{bottom!r}) (0,300) Label "Pt 4"\n'
        '  CoordSys Earth Projection 3, 33, "m", 3, 46.5, 44, 49, 700000, 6600000\n  Units "m"\n')
    got, left, top, res, crs = read_jp2(p)
    assert np.array_equal(got[0], img[:, :, 0]) and abs(left - L) < 1e-6 and abs(top - T) < 1e-6 and abs(res - RES) < 1e-9
    assert crs == "EPSG:2154"
    open(str(tmp_path / "tab.tab"), "w").write('Definition Table\n  (0,0) (0,0) Label "a",\n  (10,1) (10,0) Label "b",\n'
                                               '  (10,-10) (10,10) Label "c"\n')
    with pytest.raises(ValueError, match="north-up"):
        read_jp2(p)


def test_open_raster_is_lazy_shared_and_decodes_into_the_upload_buffer(tmp_path, monkeypatch):
    """rasterio.open() semantics: geometry without touching pixels (slicing / init_outputs only need bounds), ONE decode
    shared by everyone holding the file, and -- on a GPU host -- decoded straight into the page-locked tensor the dataset
    uploads from (emulated here with an ordinary tensor: the control flow is the same)."""
    import torch
    from flair_for_aigle_b200 import raster_io as rio
    from flair_for_aigle_b200.flair_zonal_detection import raster as raster_mod
    rng = np.random.default_rng(8)
    arr = rng.integers(0, 256, (4, 520, 700), dtype=np.uint8)
    p = str(tmp_path / "ortho.tif")
    rio.write_geotiff(p, arr, L, T, RES, epsg=2154, pixel_interleave=True, predictor=2)
    calls = []
    real_read = rio.read_window
    monkeypatch.setattr(rio, "read_window", lambda *a, **k: (calls.append(1), real_read(*a, **k))[1])

    def fake_pinned(shape, dtype, holder):
        holder["tensor"] = torch.empty(tuple(shape), dtype=torch.uint8)
        holder["array"] = holder["tensor"].numpy()
        return holder["array"]
    monkeypatch.setattr(raster_mod, "_pinned_array", fake_pinned)
    r = open_raster(p)
    assert (r.count, r.height, r.width, r.shape, r.res, r.crs) == (4, 520, 700, (520, 700), (RES, RES), "EPSG:2154")
    assert r.bounds.left == L and r.profile["dtype"] == "uint8" and not r.loaded and calls == []      # no pixel decoded yet
    assert open_raster(p) is r                                      # same file while somebody holds it
    assert np.array_equal(r.read([4, 1]), arr[[3, 0]]) and r.loaded and calls == [1]
    assert np.array_equal(open_raster(p).read(), arr) and calls == [1]                                # decoded once
    assert r.pinned_tensor is not None and r.pinned_tensor.numpy().ctypes.data == r.array.ctypes.data
    # the dataset hands exactly that tensor to the uploader (flair_zonal_detection/dataset.py: host_raster)
    from flair_for_aigle_b200.flair_zonal_detection.dataset import MultiModalSlicedDataset
    ds = MultiModalSlicedDataset.__new__(MultiModalSlicedDataset)
    ds.readers, ds.modalities, ds._device_rasters = {"AERIAL_RGBI": r}, {"AERIAL_RGBI": {"channels": [1, 2, 3, 4]}}, {}
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: (_ for _ in ()).throw(AssertionError("copied")))
    assert ds.host_raster("AERIAL_RGBI") is r.pinned_tensor
    # a rewritten file is a different raster
    rio.write_geotiff(p, arr[:, ::-1].copy(), L, T, RES, epsg=2154)
    r2 = open_raster(p)
    assert r2 is not r and np.array_equal(r2.read(), arr[:, ::-1])


def test_progressive_decode_bottom_up_behind_the_upload(tmp_path, monkeypatch):
    """ProgressiveLoad: the file decodes on a background thread in slabs of whole block rows, bottom rows first;
    ``wait_rows(lo)`` returns exactly when rows >= lo are valid (what run_streamed calls before uploading them); the result
    is the one-shot decode; a decode error reaches the waiter; JPEG 2000 files are not streamed."""
    import threading
    import torch
    from flair_for_aigle_b200 import raster_io as rio
    from flair_for_aigle_b200.flair_zonal_detection import raster as raster_mod
    rng = np.random.default_rng(9)
    arr = rng.integers(0, 256, (4, 1500, 640), dtype=np.uint8)
    p = str(tmp_path / "ortho.tif")
    rio.write_geotiff(p, arr, L, T, RES, epsg=2154, pixel_interleave=True, block=128)

    def fake_pinned(shape, dtype, holder):
        holder["tensor"] = torch.zeros(tuple(shape), dtype=torch.uint8)
        holder["array"] = holder["tensor"].numpy()
        return holder["array"]
    monkeypatch.setattr(raster_mod, "_pinned_array", fake_pinned)
    gate = threading.Semaphore(0)
    windows = []
    real_read = rio.read_window

    def gated_read(path, row0, col0, h, w, **kw):
        gate.acquire()                                   # the test lets one slab through at a time
        windows.append((row0, h))
        return real_read(path, row0, col0, h, w, **kw)
    monkeypatch.setattr(rio, "read_window", gated_read)
    monkeypatch.setattr(raster_mod.ProgressiveLoad.__init__, "__defaults__", (512, 0, None))      # slab_rows: 3 slabs of 476 / 512 / 512
    r = open_raster(p)
    prog = r.begin_progressive()
    assert prog is not None and prog.slab == 512 and r.begin_progressive() is prog and prog.lo == 1500 and not r.loaded
    gate.release()
    prog.wait_rows(1024)                                 # the bottom slab [1024, 1500)
    assert prog.lo == 1024 and windows == [(1024, 476)]
    assert np.array_equal(prog.array[:, 1024:], arr[:, 1024:]) and not prog.array[:, :1024].any()
    waiter = threading.Thread(target=prog.wait_rows, args=(600,))
    waiter.start()
    waiter.join(0.2)
    assert waiter.is_alive()                             # rows 600.. are not there yet
    gate.release()
    waiter.join(10)
    assert not waiter.is_alive() and prog.lo == 512
    for _ in range(2):
        gate.release()
    assert np.array_equal(r.read(), arr) and r.loaded    # .array waits for the rest
    assert windows == [(1024, 476), (512, 512), (0, 512)][:len(windows)] and len(windows) == 3
    assert r.pinned_tensor is prog.tensor and r.begin_progressive() is None
    # the dataset streams it: host_raster returns the tensor at once, host_rows_ready is the wait function
    monkeypatch.setattr(rio, "read_window", real_read)
    p2 = str(tmp_path / "ortho2.tif")
    rio.write_geotiff(p2, arr, L, T, RES, epsg=2154, pixel_interleave=True, block=128)
    from flair_for_aigle_b200.flair_zonal_detection.dataset import MultiModalSlicedDataset
    ds = MultiModalSlicedDataset.__new__(MultiModalSlicedDataset)
    r2 = open_raster(p2)
    ds.readers, ds.modalities, ds._device_rasters, ds._rows_ready = {"M": r2}, {"M": {"channels": None}}, {}, {}
    host = ds.host_raster("M")
    wait = ds.host_rows_ready("M")
    assert wait is not None and host is r2.begin_progressive().tensor
    wait(0, 1500)
    assert np.array_equal(host.numpy(), arr)
    # errors travel to the waiter
    p3 = str(tmp_path / "trunc.tif")
    rio.write_geotiff(p3, arr, L, T, RES, compression="none", block=128)
    raw = open(p3, "rb").read()
    open(p3, "wb").write(raw[:len(raw) // 3])
    r3 = open_raster(p3)
    with pytest.raises(rio.RasterIOError, match="outside the file"):
        r3.begin_progressive().wait_all()
    with pytest.raises(rio.RasterIOError):
        r3.read()


def test_openjpeg_driven_directly_row_windows_threads_and_fallback(tmp_path, monkeypatch):
    """flair_for_aigle_b200/openjpeg.py binds the libopenjp2 inside Pillow's wheel: row-window decodes (strips of a sharded
    zone, progressive slabs) and OpenJPEG's worker threads.  Pinned against Pillow's own decode of the same files (lossless
    AND lossy: same library, so the lossy pixels must be identical too); a JP2 zone decodes progressively and strip-wise like a
    GeoTIFF; without the library the readers fall back to Pillow."""
    from PIL import Image, features
    if not features.check("jpg_2000"):
        pytest.skip("Pillow without OpenJPEG")
    from flair_for_aigle_b200 import openjpeg as oj
    from flair_for_aigle_b200.flair_zonal_detection import geotiff, raster as raster_mod
    rng = np.random.default_rng(10)
    img = np.kron(rng.integers(0, 256, (40, 30, 4)).astype(np.uint8), np.ones((32, 32, 1), np.uint8))     # 1280 x 960 x 4
    img[::7, ::5] += 3
    want = img.transpose(2, 0, 1)
    lossless, tiled, lossy, gray = (str(tmp_path / n) for n in ("a.jp2", "t.jp2", "l.jp2", "g.j2k"))
    Image.fromarray(img).save(lossless, format="JPEG2000", irreversible=False)
    Image.fromarray(img).save(tiled, format="JPEG2000", irreversible=False, tile_size=(256, 256))
    Image.fromarray(img[:, :, :3]).save(lossy, format="JPEG2000", irreversible=True, quality_mode="rates", quality_layers=[20])
    Image.fromarray(img[:, :, 0]).save(gray, format="JPEG2000", irreversible=False)                         # bare codestream
    assert oj.info(lossless) == oj.JP2Info(960, 1280, 4, 960, 1280, oj.info(lossless).threads_supported)
    assert oj.info(tiled)[:5] == (960, 1280, 4, 256, 256) and oj.info(gray)[:3] == (960, 1280, 1)
    for path, ref in ((lossless, want), (tiled, want), (gray, want[:1])):
        for threads in (1, 0):
            assert np.array_equal(oj.read_rows(path, threads=threads), ref)
        for r0, r1 in ((0, 1), (0, 256), (255, 257), (700, 1280), (1279, 1280)):
            assert np.array_equal(oj.read_rows(path, r0, r1), ref[:, r0:r1]), (path, r0, r1)
    with Image.open(lossy) as im:
        pil = np.asarray(im).transpose(2, 0, 1)
    assert np.array_equal(oj.read_rows(lossy), pil) and not np.array_equal(pil, want[:3])      # lossy, yet the same decoder
    big = np.empty((4, 1400, 960), np.uint8)
    oj.read_rows(tiled, 100, 1000, out=big[:, 200:1100])                       # into a row-strided slab of a larger buffer
    assert np.array_equal(big[:, 200:1100], want[:, 100:1000])
    with pytest.raises(oj.OpenJPEGError, match="outside"):
        oj.read_rows(lossless, 10, 2000)
    bad = tmp_path / "bad.jp2"
    bad.write_bytes(b"\x00" * 100)
    with pytest.raises(oj.OpenJPEGError, match="neither"):
        oj.info(str(bad))
    sixteen = str(tmp_path / "s.jp2")
    Image.fromarray((img[:, :, 0].astype(np.uint16) * 200)).save(sixteen, format="JPEG2000", irreversible=False)
    with pytest.raises(oj.OpenJPEGError, match="16-bit"):
        oj.info(sixteen)
    # a georeferenced JP2 zone: progressive (slabs on tile rows) and strip-wise, like a GeoTIFF
    zone = str(tmp_path / "zone.jp2")
    Image.fromarray(img).save(zone, format="JPEG2000", irreversible=False, tile_size=(256, 256))
    raw = open(zone, "rb").read()
    at = raw.index(b"jp2c") - 4
    open(zone, "wb").write(raw[:at] + _geojp2_box(L, T, RES, 2154) + raw[at:])
    monkeypatch.setattr(raster_mod.ProgressiveLoad.__init__, "__defaults__", (512, 0, None))
    r = open_raster(zone)
    prog = r.begin_progressive()
    assert prog is not None and prog.slab == 512 and not r.loaded
    prog.wait_rows(1024)
    assert np.array_equal(prog.array[:, 1024:], want[:, 1024:])
    assert np.array_equal(r.read(), want)
    del r, prog
    import gc
    gc.collect()
    strip = open_raster(zone).row_strip(300, 900)                             # nobody holds the decoded zone: rows only
    assert not strip.loaded and strip.top == T - 300 * RES and np.array_equal(strip.read(), want[:, 300:900])
    # no library -> Pillow decodes (one thread, whole image), nothing is streamed
    monkeypatch.setattr(oj, "_lib", None)
    monkeypatch.setattr(oj, "lib", lambda: (_ for _ in ()).throw(oj.OpenJPEGUnavailable("test")))
    assert geotiff.row_source(zone) is None
    got, left, top, res, crs = geotiff.read_jp2(zone)
    assert np.array_equal(got, want) and (left, top, res, crs) == (L, T, RES, "EPSG:2154")

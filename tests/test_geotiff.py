"""GeoTIFF I/O at both ends of the path (flair_zonal_detection/geotiff.py) -- host code, no GPU."""
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

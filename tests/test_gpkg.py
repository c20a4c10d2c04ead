"""CPU: GeoPackage files on either side of the path (flair_zonal_detection/gpkg.py) -- the tile grid of slicing.py:116-119 and
the polygon frame of scripts/run_fast_aigle_segmentation.py:119-132 (``to_file(driver="GPKG")`` / ``gpd.read_file`` /
``pd.concat``).  No GDAL here: the files are checked against the GeoPackage specification's structural requirements with
sqlite3 and their geometry blobs with an independent little WKB parser."""
import sqlite3
import struct

import numpy as np
import pytest

from flair_for_aigle_b200.flair_zonal_detection.gpkg import read_gpkg, write_gpkg
from flair_for_aigle_b200.flair_zonal_detection.polygonize import PolygonTable

L, T, RES = 700000.0, 6600000.0, 0.2


def _square(x0, y0, s):
    return np.asarray([(x0, y0), (x0 + s, y0), (x0 + s, y0 + s), (x0, y0 + s), (x0, y0)], dtype=np.float64)


def _table():
    g = [{"type": "Polygon", "coordinates": [_square(L, T - 10, 10).tolist(), _square(L + 2, T - 8, 3).tolist()]},      # a hole
         {"type": "Polygon", "coordinates": [_square(L + 20, T - 30, 4).tolist()]},
         {"type": "Polygon", "coordinates": [(_square(L + 50.25, T - 7.5, 1.5)).tolist()]}]
    return PolygonTable(np.asarray([6, 6, 12]), np.asarray([100.0 - 9.0, 16.0, 2.25]), g, "EPSG:2154")


def _wkb_polygon(blob):
    """Independent parse of a GeoPackageBinary polygon: -> (srs_id, envelope, rings)."""
    assert blob[:2] == b"GP" and blob[2] == 0
    flags = blob[3]
    assert flags & 1 and (flags >> 1) & 7 == 1 and not flags & 0x10 and not flags & 0x20      # LE, xy envelope, not empty, standard
    srs_id, = struct.unpack("<i", blob[4:8])
    env = struct.unpack("<4d", blob[8:40])
    order, gtype, nrings = struct.unpack("<BII", blob[40:49])
    assert order == 1 and gtype == 3
    pos, rings = 49, []
    for _ in range(nrings):
        n, = struct.unpack("<I", blob[pos:pos + 4])
        rings.append(np.frombuffer(blob[pos + 4:pos + 4 + 16 * n], "<f8").reshape(n, 2))
        pos += 4 + 16 * n
    assert pos == len(blob)
    return srs_id, env, rings


def test_polygon_table_to_file_is_a_valid_geopackage(tmp_path):
    t = _table()
    p = t.to_file(str(tmp_path / "ortho_1.gpkg"), driver="GPKG")
    con = sqlite3.connect(p)
    cur = con.cursor()
    assert cur.execute("PRAGMA application_id").fetchone()[0] == 0x47504B47             # requirement 2
    assert cur.execute("PRAGMA user_version").fetchone()[0] == 10200
    assert cur.execute("PRAGMA integrity_check").fetchone()[0] == "ok"                  # requirement 6
    assert cur.execute("PRAGMA foreign_key_check").fetchall() == []                     # requirement 7
    srs = {r[0]: r[1:] for r in cur.execute("SELECT srs_id, organization, organization_coordsys_id, definition FROM gpkg_spatial_ref_sys")}
    assert {-1, 0, 4326, 2154} == set(srs) and srs[2154][:2] == ("EPSG", 2154) and srs[4326][2].startswith("GEOGCS")   # req. 11
    name, dtype, minx, miny, maxx, maxy, srs_id = cur.execute(
        "SELECT table_name, data_type, min_x, min_y, max_x, max_y, srs_id FROM gpkg_contents").fetchone()
    assert (name, dtype, srs_id) == ("ortho_1", "features", 2154)
    assert (minx, miny, maxx, maxy) == (L, T - 30, L + 51.75, T)
    assert cur.execute("SELECT * FROM gpkg_geometry_columns").fetchall() == [("ortho_1", "geom", "POLYGON", 2154, 0, 0)]
    cols = {r[1]: (r[2], r[5]) for r in cur.execute('PRAGMA table_info("ortho_1")')}
    assert cols == {"fid": ("INTEGER", 1), "geom": ("POLYGON", 0), "class_id": ("INTEGER", 0)}      # integer primary key (req. 29)
    rows = cur.execute('SELECT fid, geom, class_id FROM "ortho_1" ORDER BY fid').fetchall()
    con.close()
    assert [r[0] for r in rows] == [1, 2, 3] and [r[2] for r in rows] == [6, 6, 12]
    for (fid, blob, _), geom in zip(rows, t.geometry):
        srs_id, env, rings = _wkb_polygon(blob)
        want = [np.asarray(r) for r in geom["coordinates"]]
        assert srs_id == 2154 and len(rings) == len(want) and all(np.array_equal(a, b) for a, b in zip(rings, want))
        assert env == (want[0][:, 0].min(), want[0][:, 0].max(), want[0][:, 1].min(), want[0][:, 1].max())


def test_read_file_and_concat_like_the_product_script(tmp_path):
    """scripts/run_fast_aigle_segmentation.py:123,131-132: one GPKG per image, read back and concatenated."""
    t = _table()
    a = t.to_file(str(tmp_path / "a.gpkg"))
    shifted = PolygonTable(t.class_id[:2] + 1, t.area[:2],
                           [{"type": "Polygon", "coordinates": [(np.asarray(r) + 100.0).tolist() for r in g["coordinates"]]}
                            for g in t.geometry[:2]], "EPSG:2154")
    b = shifted.to_file(str(tmp_path / "b.gpkg"))
    back = PolygonTable.read_file(a)
    assert back.crs == "EPSG:2154" and back.class_id.tolist() == [6, 6, 12] and np.allclose(back.area, t.area)
    assert [g for g in back.geometry] == [g for g in t.geometry]
    both = PolygonTable.concat([back, PolygonTable.read_file(b), PolygonTable(np.zeros(0, np.int64), np.zeros(0), [], None)])
    assert len(both) == 5 and both.class_id.tolist() == [6, 6, 12, 7, 7] and both.crs == "EPSG:2154"
    assert both.geometry[3]["coordinates"][1][0] == [L + 102.0, T - 8 + 100.0]
    # rewriting replaces the file (geopandas' mode 'w'); GeoJSON stays available; unknown drivers are refused
    both.to_file(a)
    assert len(PolygonTable.read_file(a)) == 5
    both.to_file(str(tmp_path / "x.geojson"), driver="GeoJSON")
    with pytest.raises(ValueError, match="driver"):
        both.to_file(str(tmp_path / "x.shp"), driver="ESRI Shapefile")
    with pytest.raises(FileNotFoundError):
        PolygonTable.read_file(str(tmp_path / "missing.gpkg"))
    empty = PolygonTable(np.zeros(0, np.int64), np.zeros(0), [], None).to_file(str(tmp_path / "empty.gpkg"))
    assert len(PolygonTable.read_file(empty)) == 0 and PolygonTable.read_file(empty).crs is None


def test_traced_rings_survive_the_file(tmp_path):
    """The flat-array geometries raster_to_polygons builds (host ring tracer) go through the same writer."""
    from flair_for_aigle_b200 import native as nv
    from flair_for_aigle_b200.flair_zonal_detection.polygonize import _Geometries
    lab = np.zeros((12, 12), np.int32)
    lab[1:9, 1:9] = 13                                    # root = first pixel index of the component
    lab[3:6, 3:6] = 0                                     # a hole
    lab[10:12, 9:12] = 10 * 12 + 9
    lab[lab == 13] = 1 * 12 + 1
    roots = np.asarray([13, 129], np.int32)
    ring_root, ring_hole, ring_off, xy = nv.trace_rings(lab, roots, 0.0)
    ring_poly = np.searchsorted(roots, ring_root)
    perm = np.lexsort((np.arange(ring_root.size), ring_hole, ring_poly))
    geoms = _Geometries(xy, ring_off, perm, np.concatenate([[0], np.cumsum(np.bincount(ring_poly, minlength=2))]))
    t = PolygonTable(np.asarray([3, 7]), np.asarray([55.0, 6.0]), geoms, "EPSG:2154")
    back = PolygonTable.read_file(t.to_file(str(tmp_path / "traced.gpkg")))
    assert [len(g["coordinates"]) for g in back.geometry] == [2, 1] and np.allclose(back.area, [55.0, 6.0])
    assert [g for g in back.geometry] == [g for g in t.geometry]


def test_tile_grid_written_like_the_reference(tmp_path):
    """slicing.py:116-119: write_dataframe -> <output_name>_slicing_job.gpkg with the frame's columns and the tile boxes."""
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    r = ZoneRaster(np.broadcast_to(np.zeros((1, 1, 1), np.uint8), (1, 700, 1000)), L, T, RES, crs="EPSG:2154", name="mem://grid")
    cfg = {"img_pixels_detection": 512, "margin": 64, "output_name": "zone", "reference_resolution": RES,
           "output_path": str(tmp_path), "write_dataframe": True}
    gdf = generate_patches_from_reference(cfg, r, None)
    cols, geoms, crs = read_gpkg(str(tmp_path / "zone_slicing_job.gpkg"))
    assert crs == "EPSG:2154" and len(geoms) == len(gdf) == 6
    assert list(cols) == [c for c in gdf.columns if c != "geometry"]
    assert cols["id"].tolist() == gdf["id"].tolist() and np.array_equal(cols["left"], gdf["left"].to_numpy())
    assert cols["job_done"].dtype.kind == "i" and cols["input_id"][0] == "mem://grid"
    for ring, g in zip(geoms, gdf.geometry):
        x0, y0, x1, y1 = g.bounds
        assert len(ring) == 1 and ring[0].tolist() == [[x1, y0], [x1, y1], [x0, y1], [x0, y0], [x1, y0]]


def test_write_gpkg_argument_checks(tmp_path):
    with pytest.raises((ValueError, IndexError)):
        write_gpkg(str(tmp_path / "bad.gpkg"), [[_square(0, 0, 1)], [_square(2, 2, 1)]], {"class_id": [1]}, "EPSG:2154")
    p = write_gpkg(str(tmp_path / "odd name.gpkg"), [[_square(0, 0, 1)]], {"score": [0.5], "label": ["roof"]}, None)
    cols, geoms, crs = read_gpkg(p)
    assert crs is None and cols["score"].tolist() == [0.5] and cols["label"].tolist() == ["roof"] and len(geoms) == 1
    with pytest.raises(ValueError, match="no feature layer"):
        read_gpkg(p, layer="nope")
    for bad in ('a"b', "fid", "GEOM", ""):
        with pytest.raises(ValueError, match="column name"):
            write_gpkg(str(tmp_path / "c.gpkg"), [[_square(0, 0, 1)]], {bad: [1]}, None)


def test_confidence_column_and_the_small_reference_helpers(tmp_path):
    """vectorize_segmentation's third column (inference.py:588,610: every polygon carries the mean confidence of ALL pixels of
    its class), carried through to_file / read_file / concat; create_polygon_from_bounds (postprocess.py:55-66)."""
    from flair_for_aigle_b200.flair_zonal_detection.polygonize import class_mean_confidence
    from flair_for_aigle_b200.flair_zonal_detection.postprocess import create_polygon_from_bounds
    rng = np.random.default_rng(3)
    labels, conf = rng.integers(0, 6, (80, 90)), rng.random((80, 90))
    ids = np.asarray([5, 1, 1, 3])
    got = class_mean_confidence(labels, conf, ids)
    assert np.allclose(got, [conf[labels == v].mean() for v in ids], rtol=1e-12)
    t = _table()
    t.confidence = np.asarray([0.25, 0.25, 0.75])
    back = PolygonTable.read_file(t.to_file(str(tmp_path / "c.gpkg")))
    assert back.confidence.tolist() == [0.25, 0.25, 0.75] and back.class_id.tolist() == [6, 6, 12]
    both = PolygonTable.concat([back, back])
    assert both.confidence.tolist() == [0.25, 0.25, 0.75] * 2
    assert PolygonTable.concat([back, _table()]).confidence is None          # a table without the column: dropped, not invented
    # shapely: box(x_min, y_max, x_max, y_min) starts at (maxx, miny) = (x_max, y_max) and runs counter-clockwise in its own frame
    assert create_polygon_from_bounds(0, 2, 1, 5) == {"type": "Polygon", "coordinates": (((2.0, 5.0), (2.0, 1.0), (0.0, 1.0),
                                                                                       (0.0, 5.0), (2.0, 5.0)),)}

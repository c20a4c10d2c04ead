"""GeoPackage (OGC 12-128r15, version 1.2) vector files on the host: what ``GeoDataFrame.to_file(path, driver="GPKG")`` and
``gpd.read_file(path)`` do for the two frames on either side of the zonal path --

  the tile grid           flair_zonal_detection/slicing.py:116-119      ``<output_name>_slicing_job.gpkg``
  the polygons            scripts/run_fast_aigle_segmentation.py:119-123 ``raster_to_polygons(...).to_file(..., driver="GPKG")``
                          and :131 (``gpd.read_file`` of every per-image file before the aggregation)

geopandas / fiona / GDAL are not in this image; a GeoPackage is a SQLite database with three metadata tables and one
feature table whose geometry column holds "GeoPackageBinary" blobs (a small header + ISO WKB), so Python's ``sqlite3``
writes and reads it.  One layer per file, polygons (with holes) only, attribute columns of integer / real / text type.  No
R-tree index is written (an optional extension; readers build their own).
"""
from __future__ import annotations

import datetime
import os
import sqlite3
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .geotiff import _epsg, _is_geographic

APPLICATION_ID = 0x47504B47            # 'GPKG'
USER_VERSION = 10200                   # 1.2.0
_WKT_4326 = ('GEOGCS["WGS 84",DATUM["WGS_1984",SPHEROID["WGS 84",6378137,298.257223563,AUTHORITY["EPSG","7030"]],'
             'AUTHORITY["EPSG","6326"]],PRIMEM["Greenwich",0,AUTHORITY["EPSG","8901"]],UNIT["degree",0.0174532925199433,'
             'AUTHORITY["EPSG","9122"]],AUTHORITY["EPSG","4326"]]')


def _blob(rings: Sequence[np.ndarray], srs_id: int) -> Tuple[bytes, Tuple[float, float, float, float]]:
    """GeoPackageBinary: 'GP', version 0, flags (little endian, envelope [minx, maxx, miny, maxy]), srs_id, envelope, then
    the polygon as little-endian ISO WKB (type 3: ring count, per ring point count + x, y doubles)."""
    ext = np.asarray(rings[0], dtype="<f8")
    minx, miny = ext.min(axis=0)
    maxx, maxy = ext.max(axis=0)
    parts = [b"GP\x00\x03", struct.pack("<i4d", srs_id, minx, maxx, miny, maxy), struct.pack("<BII", 1, 3, len(rings))]
    for r in rings:
        a = np.ascontiguousarray(r, dtype="<f8")
        parts.append(struct.pack("<I", a.shape[0]))
        parts.append(a.tobytes())
    return b"".join(parts), (float(minx), float(miny), float(maxx), float(maxy))


def _sql_type(values: np.ndarray) -> str:
    if values.dtype.kind in "iub":
        return "INTEGER"
    if values.dtype.kind == "f":
        return "REAL"
    return "TEXT"


def write_gpkg(path: str, geometries: Iterable[Sequence[np.ndarray]], columns: Dict[str, Sequence], crs: Optional[str] = None,
               layer: Optional[str] = None) -> str:
    """One polygon layer.  ``geometries``: per feature the list of closed rings ([n,2] arrays; exterior first, then holes);
    ``columns``: attribute name -> one value per feature.  An existing file is replaced (geopandas' default mode 'w')."""
    layer = layer or os.path.splitext(os.path.basename(path))[0]
    if not layer.replace("_", "").replace("-", "").isalnum():
        layer = "layer"
    epsg = _epsg(crs)
    srs_id = epsg if epsg is not None else -1
    cols = {k: np.asarray(v) for k, v in columns.items()}
    for k in cols:
        if not isinstance(k, str) or not k or '"' in k or k.lower() in ("fid", "geom"):
            raise ValueError(f"column name {k!r} cannot be used in a GeoPackage feature table")
    if os.path.exists(path):
        os.remove(path)
    con = sqlite3.connect(path)
    try:
        cur = con.cursor()
        cur.execute(f"PRAGMA application_id = {APPLICATION_ID}")
        cur.execute(f"PRAGMA user_version = {USER_VERSION}")
        cur.executescript("""
            CREATE TABLE gpkg_spatial_ref_sys (srs_name TEXT NOT NULL, srs_id INTEGER NOT NULL PRIMARY KEY,
                organization TEXT NOT NULL, organization_coordsys_id INTEGER NOT NULL, definition TEXT NOT NULL, description TEXT);
            CREATE TABLE gpkg_contents (table_name TEXT NOT NULL PRIMARY KEY, data_type TEXT NOT NULL, identifier TEXT UNIQUE,
                description TEXT DEFAULT '', last_change DATETIME NOT NULL DEFAULT (strftime('%Y-%m-%dT%H:%M:%fZ','now')),
                min_x DOUBLE, min_y DOUBLE, max_x DOUBLE, max_y DOUBLE, srs_id INTEGER,
                CONSTRAINT fk_gc_r_srs_id FOREIGN KEY (srs_id) REFERENCES gpkg_spatial_ref_sys(srs_id));
            CREATE TABLE gpkg_geometry_columns (table_name TEXT NOT NULL, column_name TEXT NOT NULL, geometry_type_name TEXT NOT NULL,
                srs_id INTEGER NOT NULL, z TINYINT NOT NULL, m TINYINT NOT NULL,
                CONSTRAINT pk_geom_cols PRIMARY KEY (table_name, column_name),
                CONSTRAINT fk_gc_tn FOREIGN KEY (table_name) REFERENCES gpkg_contents(table_name),
                CONSTRAINT fk_gc_srs FOREIGN KEY (srs_id) REFERENCES gpkg_spatial_ref_sys(srs_id));
        """)
        srs_rows = [("Undefined cartesian SRS", -1, "NONE", -1, "undefined", "undefined cartesian coordinate reference system"),
                    ("Undefined geographic SRS", 0, "NONE", 0, "undefined", "undefined geographic coordinate reference system"),
                    ("WGS 84 geodetic", 4326, "EPSG", 4326, _WKT_4326, "longitude/latitude coordinates in decimal degrees on the WGS 84 spheroid")]
        if epsg is not None and epsg != 4326:
            # the definition of any other code is left to the reader's EPSG registry (organization + code identify it)
            srs_rows.append((f"EPSG:{epsg}", epsg, "EPSG", epsg, "undefined",
                             "geographic" if _is_geographic(crs, epsg) else "projected"))
        cur.executemany("INSERT INTO gpkg_spatial_ref_sys VALUES (?,?,?,?,?,?)", srs_rows)
        decl = "".join(f', "{k}" {_sql_type(v)}' for k, v in cols.items())
        cur.execute(f'CREATE TABLE "{layer}" (fid INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, geom POLYGON{decl})')
        names = ", ".join(f'"{k}"' for k in cols)
        marks = ", ".join("?" for _ in range(len(cols) + 1))
        lists = [v.tolist() for v in cols.values()]
        box = [np.inf, np.inf, -np.inf, -np.inf]

        def rows():
            for i, rings in enumerate(geometries):
                blob, (x0, y0, x1, y1) = _blob(rings, srs_id)
                box[0], box[1], box[2], box[3] = min(box[0], x0), min(box[1], y0), max(box[2], x1), max(box[3], y1)
                yield (blob, *[col[i] for col in lists])
        cur.executemany(f'INSERT INTO "{layer}" (geom{", " + names if names else ""}) VALUES ({marks})', rows())
        n = cur.execute(f'SELECT COUNT(*) FROM "{layer}"').fetchone()[0]
        for k, v in cols.items():
            if len(v) != n:
                raise ValueError(f"column '{k}' has {len(v)} values for {n} geometries")
        bounds = [None] * 4 if n == 0 else [float(b) for b in box]
        now = datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%S.%f")[:-3] + "Z"
        cur.execute("INSERT INTO gpkg_contents (table_name, data_type, identifier, description, last_change, min_x, min_y, max_x, "
                    "max_y, srs_id) VALUES (?,?,?,?,?,?,?,?,?,?)", (layer, "features", layer, "", now, *bounds, srs_id))
        cur.execute("INSERT INTO gpkg_geometry_columns VALUES (?,?,?,?,?,?)", (layer, "geom", "POLYGON", srs_id, 0, 0))
        con.commit()
    finally:
        con.close()
    return path


def _parse_blob(blob: bytes) -> List[np.ndarray]:
    if blob[:2] != b"GP":
        raise ValueError("not a GeoPackageBinary geometry")
    flags = blob[3]
    env = (flags >> 1) & 7
    pos = 8 + {0: 0, 1: 32, 2: 48, 3: 48, 4: 64}[env]
    if flags & 0x10:
        return []                                    # empty geometry
    e = "<" if blob[pos] == 1 else ">"
    gtype = struct.unpack(e + "I", blob[pos + 1:pos + 5])[0]
    if gtype % 1000 != 3:
        raise NotImplementedError(f"WKB geometry type {gtype} (only polygons are read)")
    dims = 2 + (1 if gtype // 1000 in (1, 3) else 0) + (1 if gtype // 1000 in (2, 3) else 0)
    nrings = struct.unpack(e + "I", blob[pos + 5:pos + 9])[0]
    pos += 9
    rings = []
    for _ in range(nrings):
        npts = struct.unpack(e + "I", blob[pos:pos + 4])[0]
        pos += 4
        a = np.frombuffer(blob, dtype=e + "f8", count=npts * dims, offset=pos).reshape(npts, dims)
        rings.append(np.array(a[:, :2], dtype=np.float64))
        pos += 8 * npts * dims
    return rings


def read_gpkg(path: str, layer: Optional[str] = None):
    """``gpd.read_file(path)`` for a polygon layer -> (columns {name: numpy array}, geometries [rings per feature], crs)."""
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    con = sqlite3.connect(path)
    try:
        cur = con.cursor()
        if cur.execute("PRAGMA application_id").fetchone()[0] != APPLICATION_ID:
            raise ValueError(f"{path}: not a GeoPackage (application_id)")
        feats = cur.execute("SELECT c.table_name, g.column_name, g.srs_id FROM gpkg_contents c JOIN gpkg_geometry_columns g "
                            "ON c.table_name = g.table_name WHERE c.data_type = 'features'").fetchall()
        if layer is not None:
            feats = [f for f in feats if f[0] == layer]
        if not feats:
            raise ValueError(f"{path}: no feature layer" + (f" named '{layer}'" if layer else ""))
        table, geom_col, srs_id = feats[0]
        org = cur.execute("SELECT organization, organization_coordsys_id FROM gpkg_spatial_ref_sys WHERE srs_id = ?",
                          (srs_id,)).fetchone()
        crs = f"EPSG:{org[1]}" if org and str(org[0]).upper() == "EPSG" else None
        info = cur.execute(f'PRAGMA table_info("{table}")').fetchall()
        pk = [c[1] for c in info if c[5]]
        attrs = [c[1] for c in info if c[1] != geom_col and c[1] not in pk]
        sel = ", ".join([f'"{geom_col}"'] + [f'"{a}"' for a in attrs])
        order = f' ORDER BY "{pk[0]}"' if pk else ""
        rows = cur.execute(f'SELECT {sel} FROM "{table}"{order}').fetchall()
    finally:
        con.close()
    geometries = [_parse_blob(r[0]) for r in rows]
    columns = {a: np.asarray([r[k + 1] for r in rows]) for k, a in enumerate(attrs)}
    return columns, geometries, crs

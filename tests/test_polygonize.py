"""Polygonisation of the class raster (inference.py:356-407): the host ring tracer of the C-ABI library on CPU against
the oracle's scipy component structure (polygons rasterised back must reproduce every component exactly), and the GPU
connected-component labelling bit for bit against the oracle."""
import numpy as np
import pytest
import torch

from oracle.polygons import component_table, label_components, rasterize_even_odd, ring_area2


def _rasters():
    rng = np.random.default_rng(2025)
    out = {}
    out["noise19"] = rng.integers(0, 19, (70, 97), dtype=np.uint8)
    out["noise3"] = rng.integers(0, 3, (64, 300), dtype=np.uint8)
    blobs = np.zeros((120, 140), np.uint8)
    yy, xx = np.mgrid[:120, :140]
    for k in range(14):
        cy, cx, r = rng.integers(0, 120), rng.integers(0, 140), rng.integers(4, 30)
        blobs[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = k % 5 + 1
    out["blobs"] = blobs
    out["uniform"] = np.full((40, 513), 6, np.uint8)
    out["checker"] = ((np.indices((33, 65)).sum(axis=0)) % 2).astype(np.uint8)
    ring = np.zeros((50, 50), np.uint8)
    ring[5:45, 5:45] = 1
    ring[10:40, 10:40] = 0
    ring[15:35, 15:35] = 1          # island inside the hole
    ring[20:30, 20:30] = 2
    out["nested"] = ring
    out["row"] = rng.integers(0, 2, (1, 77), dtype=np.uint8)
    out["col"] = rng.integers(0, 2, (77, 1), dtype=np.uint8)
    spiral = np.zeros((41, 41), np.uint8)
    for k in range(0, 20, 2):
        spiral[k, k:41 - k] = 1
        spiral[k:41 - k, 40 - k] = 1
        spiral[40 - k, k:41 - k] = 1
        spiral[k + 2:41 - k, k] = 1
    out["spiral"] = spiral
    return out


RASTERS = _rasters()


@pytest.mark.parametrize("name", sorted(RASTERS))
def test_ring_tracer_reproduces_every_component(name):
    from flair_for_aigle_b200 import native as nv
    raster = RASTERS[name]
    H, W = raster.shape
    labels = label_components(raster)
    roots, areas, _ = component_table(raster, labels)
    ring_root, ring_hole, off, xy = nv.trace_rings(labels, roots, 0.0)
    assert set(ring_root.tolist()) == set(roots.tolist())
    by_root = {}
    for k in range(ring_root.size):
        by_root.setdefault(int(ring_root[k]), []).append((bool(ring_hole[k]), xy[off[k]:off[k + 1]]))
    for r, a in zip(roots.tolist(), areas.tolist()):
        rings = by_root[r]
        ext = [g for h, g in rings if not h]
        assert len(ext) == 1, "one exterior ring per 4-connected component"
        for h, g in rings:
            assert np.array_equal(g[0], g[-1]) and len(g) >= 5
            assert (ring_area2(g) < 0) == h
            d = np.abs(np.diff(g, axis=0))
            assert np.all((d[:, 0] == 0) != (d[:, 1] == 0)), "rectilinear, no zero-length and no collinear duplicate edges"
        assert abs(sum(ring_area2(g) for _, g in rings) / 2 - a) < 1e-9          # shoelace area = pixel count
        assert np.array_equal(rasterize_even_odd([g for _, g in rings], H, W), labels == r)


def test_ring_tracer_keep_list_and_simplification():
    from flair_for_aigle_b200 import native as nv
    raster = RASTERS["blobs"]
    labels = label_components(raster)
    roots, areas, classes = component_table(raster, labels)
    keep = roots[(classes != 0) & (areas >= 30)]
    rr, hole, off, xy = nv.trace_rings(labels, keep, 0.0)
    assert set(rr.tolist()) == set(keep.tolist())
    rr2, hole2, off2, xy2 = nv.trace_rings(labels, keep, 0.5)         # the reference's 0.1 m at 0.2 m / px
    assert np.array_equal(rr, rr2) and np.array_equal(hole, hole2)
    assert xy2.shape[0] < xy.shape[0]
    for k in range(rr.size):
        full, simp = xy[off[k]:off[k + 1]], xy2[off2[k]:off2[k + 1]]
        assert np.array_equal(simp[0], simp[-1]) and len(simp) >= 4
        assert {tuple(p) for p in simp} <= {tuple(p) for p in full}
        # every dropped corner lies within the tolerance of the simplified ring
        a, b = simp[:-1], simp[1:]
        for p in full:
            ab, ap = b - a, p - a
            t = np.clip((ap * ab).sum(1) / np.maximum((ab * ab).sum(1), 1e-30), 0, 1)
            d = np.sqrt(((a + t[:, None] * ab - p) ** 2).sum(1)).min()
            assert d <= 0.5 + 1e-9
    # an empty keep list traces nothing
    rr3, _, off3, xy3 = nv.trace_rings(labels, np.zeros(0, np.int32), 0.0)
    assert rr3.size == 0 and xy3.shape == (0, 2) and off3.tolist() == [0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(RASTERS))
def test_ccl_labels_are_bit_exact(cuda, name):
    from flair_for_aigle_b200 import native as nv
    raster = RASTERS[name]
    want = label_components(raster)
    t = torch.from_numpy(raster).to(cuda)
    labels = nv.ccl_label(t)
    assert np.array_equal(labels.cpu().numpy(), want)
    roots, areas, classes = nv.ccl_components(t, labels)
    wr, wa, wc = component_table(raster, want)
    assert np.array_equal(roots, wr) and np.array_equal(areas, wa) and np.array_equal(classes, wc)


@pytest.mark.gpu
def test_ccl_on_a_large_smooth_class_map(cuda):
    """2048 x 3000 px, class map with large regions (long union chains) and speckle: labels identical to the oracle."""
    from flair_for_aigle_b200 import native as nv
    from flair_for_aigle_b200.synthetic import synthetic_raster
    arr = synthetic_raster(2048, 3000, seed=3)
    raster = ((arr[0].astype(np.int32) + arr[1]) // 40 % 19).astype(np.uint8)
    want = label_components(raster)
    labels = nv.ccl_label(torch.from_numpy(raster).to(cuda))
    assert np.array_equal(labels.cpu().numpy(), want)
    again = nv.ccl_label(torch.from_numpy(raster).to(cuda))
    assert torch.equal(labels, again)                                # order independent


@pytest.mark.gpu
def test_raster_to_polygons_end_to_end(cuda):
    """raster_to_polygons (inference.py:375-407) on a class raster: classes / areas / filters as the reference's rules,
    and the polygons rasterise back to exactly the pixels of the kept components."""
    from flair_for_aigle_b200.flair_zonal_detection.inference import raster_to_polygons
    rng = np.random.default_rng(5)
    raster = RASTERS["blobs"].copy()
    raster[raster == 0] = 18                                         # background
    raster[rng.random(raster.shape) < 0.02] = 7                      # speckle below min_area
    H, W = raster.shape
    left, top, res = 700000.0, 6600000.0, 0.2
    table = raster_to_polygons((raster, left, top, res, "EPSG:2154"), min_area=1.0, simplification=0.0, device=cuda)
    labels = label_components(raster)
    roots, areas, classes = component_table(raster, labels)
    keep = (classes != 18) & (areas * res * res >= 1.0)
    assert len(table) == int(keep.sum()) > 0
    assert sorted(table.class_id.tolist()) == sorted(classes[keep].tolist())
    assert np.allclose(np.sort(table.area), np.sort(areas[keep] * res * res))
    assert list(table.class_id) == sorted(table.class_id)            # ordered by class
    covered = np.zeros((H, W), bool)
    for row in table:
        rings = [[((x - left) / res, (top - y) / res) for x, y in ring] for ring in row["geometry"]["coordinates"]]
        m = rasterize_even_odd(rings, H, W)
        assert (raster[m] == row["class_id"]).all() and not (covered & m).any()
        covered |= m
    assert np.array_equal(covered, np.isin(labels, roots[keep]))
    fc = table.to_geojson()
    assert fc["features"][0]["geometry"]["type"] == "Polygon" and fc["crs"]["properties"]["name"] == "EPSG:2154"
    simplified = raster_to_polygons((raster, left, top, res), simplification=0.1, device=cuda)
    assert len(simplified) == len(table)
    n_full = sum(len(r) for g in table.geometry for r in g["coordinates"])
    n_simp = sum(len(r) for g in simplified.geometry for r in g["coordinates"])
    assert n_simp < n_full


def test_ring_tracer_threads_give_the_sequential_result(monkeypatch):
    """The host tracer deals the components to threads (every ring is walked once, by the owner of its component) and merges
    the rings by their start position, so any thread count returns exactly the single-thread arrays (order included)."""
    from flair_for_aigle_b200 import native as nv
    rng = np.random.default_rng(77)
    H, W = 700, 420
    raster = rng.integers(0, 4, (H // 10, W // 10), dtype=np.uint8).repeat(10, axis=0).repeat(10, axis=1)
    raster[rng.random((H, W)) < 0.03] = 5                                      # speckle: many small holes
    raster[100:600, 200:210] = 6                                               # a tall component crossing every band
    labels = label_components(raster)
    roots, areas, classes = component_table(raster, labels)
    keep = roots[areas >= 3]
    monkeypatch.setenv("FZ_TRACE_THREADS", "1")
    ref = nv.trace_rings(labels, keep, 0.5)
    for t in ("2", "5", "10"):
        monkeypatch.setenv("FZ_TRACE_THREADS", t)
        got = nv.trace_rings(labels, keep, 0.5)
        assert all(np.array_equal(a, b) for a, b in zip(ref, got)), t
    assert ref[0].size > 1000

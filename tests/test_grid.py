"""CPU: tile grid / window arithmetic.  Oracle (loop restatement of slicing.py:51-112) against
the known answers of SURVEY.md H7, and the product's separable implementation against the
oracle, bit-exactly."""
import numpy as np
import pytest

from oracle.grid import Georef, generate_patches, tile_plan as oracle_plan
from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster
from flair_for_aigle_b200.flair_zonal_detection.slicing import (generate_patches_from_reference, ownership_windows,
                                                                tile_plan)

L, T, RES = 700000.0, 6600000.0, 0.2

KNOWN = [  # (W, H, margin, n_tiles)   patch 512, res 0.2  -- SURVEY.md H7
    (10000, 10000, 64, 729), (10000, 10000, 128, 1600), (10000, 10000, 40, 576), (20000, 20000, 64, 2809),
    (1000, 700, 64, 6), (2048, 2048, 128, 64), (777, 1300, 40, 8),
]


@pytest.mark.parametrize("W,H,margin,n", KNOWN)
def test_oracle_known_counts(W, H, margin, n):
    geo = Georef(L, T, RES, W, H)
    tiles = generate_patches(512, margin, RES, geo)
    assert len(tiles) == n
    plan = oracle_plan(tiles, geo, 512, margin)
    cov = np.zeros((H, W), np.uint8)
    for r in plan:
        cov[r[2]:r[2] + r[4], r[3]:r[3] + r[5]] += 1
    assert cov.min() >= 1  # every pixel is covered


def test_oracle_known_offsets():
    geo = Georef(L, T, RES, 1000, 700)
    tiles = generate_patches(512, 64, RES, geo)
    plan = oracle_plan(tiles, geo, 512, 64)
    assert sorted(set(plan[:, 3].tolist())) == [0, 384, 616]   # inner lefts
    assert sorted(set(plan[:, 2].tolist())) == [0, 316]        # inner tops
    assert sorted(set(plan[:, 1].tolist())) == [-64, 320, 552]  # read-window column offsets
    # x-outer ascending, y-inner bottom -> top
    assert plan[0, 3] == 0 and plan[0, 2] == 316 and plan[1, 2] == 0
    assert tiles[0]["id"].startswith("1-")


def test_oracle_60k_count():
    geo = Georef(L, T, RES, 60000, 60000)
    assert len(generate_patches(512, 64, RES, geo)) == 24649


def _product_tiles(W, H, P, margin, res=RES, geozone=None, left=L, top=T):
    r = ZoneRaster(np.zeros((1, 1, 1), np.uint8), left, top, res)
    r.array = np.broadcast_to(r.array, (1, H, W))
    cfg = {"img_pixels_detection": P, "margin": margin, "output_path": ".", "output_name": "t",
           "reference_modality": "AERIAL_RGBI", "reference_resolution": res}
    return r, generate_patches_from_reference(cfg, r, geozone)


CASES = KNOWN[:3] + KNOWN[4:] + [(700, 512, 0, None), (512, 512, 100, None), (1500, 90, 31, None),
                                 (5000, 3000, 17, None)]


@pytest.mark.parametrize("W,H,margin,n", CASES)
def test_product_grid_equals_oracle(W, H, margin, n):
    geo = Georef(L, T, RES, W, H)
    ref = generate_patches(512, margin, RES, geo)
    r, gdf = _product_tiles(W, H, 512, margin)
    assert len(gdf) == len(ref)
    if n is not None:
        assert len(gdf) == n
    for i, t in enumerate(ref):
        row = gdf.iloc[i]
        assert row["id"] == t["id"]
        for k in ("left", "bottom", "right", "top", "left_o", "bottom_o", "right_o", "top_o"):
            assert row[k] == t[k], (i, k)           # bit-exact float64
        assert row["geometry"].bounds == t["geometry"]
    b = r.bounds
    plan = tile_plan(gdf, {"left": b.left, "bottom": b.bottom, "right": b.right, "top": b.top}, RES, 512, margin)
    assert np.array_equal(plan, oracle_plan(ref, geo, 512, margin))


def test_product_grid_random_shapes():
    rng = np.random.default_rng(0)
    for _ in range(40):
        P = int(rng.choice([64, 128, 256]))
        margin = int(rng.integers(0, P // 2 - 1))
        W, H = int(rng.integers(1, 900)), int(rng.integers(1, 900))
        res = float(rng.choice([0.2, 0.5, 1.0, 0.15]))
        left, top = float(rng.integers(0, 10 ** 6)) + 0.3, float(rng.integers(10 ** 6, 7 * 10 ** 6)) + 0.7
        geo = Georef(left, top, res, W, H)
        ref = generate_patches(P, margin, res, geo)
        r, gdf = _product_tiles(W, H, P, margin, res, None, left, top)
        assert len(gdf) == len(ref)
        for i, t in enumerate(ref):
            row = gdf.iloc[i]
            assert (row["left"], row["bottom"], row["right"], row["top"]) == (t["left"], t["bottom"], t["right"], t["top"])
            assert row["id"] == t["id"]


def test_geozone_crop_and_miss():
    geo = Georef(L, T, RES, 4000, 3000)
    bbox = (L + 100.03, T - 500.0, L + 433.3, T - 77.7)
    ref = generate_patches(512, 64, RES, geo, geozone_bbox=bbox)
    _, gdf = _product_tiles(4000, 3000, 512, 64, geozone=bbox)
    assert len(ref) == len(gdf) > 0
    assert [t["id"] for t in ref] == list(gdf["id"])
    assert [t["left"] for t in ref] == list(gdf["left"])
    miss = (L - 500.0, T + 10.0, L - 100.0, T + 200.0)
    assert generate_patches(512, 64, RES, geo, geozone_bbox=miss) == []
    _, empty = _product_tiles(4000, 3000, 512, 64, geozone=miss)
    assert len(empty) == 0


@pytest.mark.parametrize("W,H,margin,P", [(1000, 700, 64, 512), (2048, 2048, 128, 512), (777, 1300, 40, 512),
                                          (300, 217, 16, 128), (140, 411, 8, 128), (256, 256, 32, 128)])
def test_ownership_is_last_writer(W, H, margin, P):
    geo = Georef(L, T, RES, W, H)
    plan = oracle_plan(generate_patches(P, margin, RES, geo), geo, P, margin)
    owner = np.full((H, W), -1, np.int32)
    for i, r in enumerate(plan):          # sequential writes, later tiles overwrite
        owner[r[2]:r[2] + r[4], r[3]:r[3] + r[5]] = i
    own = ownership_windows(plan)
    painted = np.full((H, W), -1, np.int32)
    for i, o in enumerate(own):
        if o[1] > o[0] and o[3] > o[2]:
            assert (painted[o[0]:o[1], o[2]:o[3]] == -1).all()   # owned windows are disjoint
            painted[o[0]:o[1], o[2]:o[3]] = i
    assert np.array_equal(painted, owner)


def test_raster_smaller_than_inner_tile_is_rejected():
    # the reference would hand rasterio a window with a negative offset (inference.py:343)
    r, gdf = _product_tiles(333, 512, 512, 0)
    b = r.bounds
    with pytest.raises(ValueError):
        tile_plan(gdf, {"left": b.left, "bottom": b.bottom, "right": b.right, "top": b.top}, RES, 512, 0)

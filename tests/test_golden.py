"""CPU: committed golden fixtures (tests/golden, made by tests/golden/make_reference_golden.py by running the
REFERENCE's own slicing / dataset / inference_and_write / convert code at the SURVEY.md H7 cases) against the oracle AND
the product's grid implementation.  These travel to the GPU box, where /root/reference does not exist."""
import json
import os

import numpy as np
import pytest

from oracle.convert import convert
from oracle.grid import Georef, generate_patches, tile_plan as oracle_plan
from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster
from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference, tile_plan

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
L, T, RES = 700000.0, 6600000.0, 0.2
PLANS = json.load(open(os.path.join(GOLD, "grid_plans.json")))
H7 = {"1000x700_m64": 6, "2048x2048_m128": 64, "777x1300_m40": 8, "10000x10000_m64": 729,
      "10000x10000_m128": 1600, "10000x10000_m40": 576, "20000x20000_m64": 2809}


@pytest.mark.parametrize("key", sorted(k for k in PLANS if not k.startswith("_")))
def test_grid_golden(key):
    g = PLANS[key]
    assert g["n_tiles"] == H7[key]                       # the survey's independently derived counts
    dims, m = key.split("_m")
    W, H = (int(v) for v in dims.split("x"))
    m = int(m)
    geo = Georef(L, T, RES, W, H)
    tiles = generate_patches(512, m, RES, geo)
    plan = oracle_plan(tiles, geo, 512, m)
    assert len(tiles) == g["n_tiles"] and [t["id"] for t in tiles[:4]] == g["first_ids"] and tiles[-1]["id"] == g["last_id"]
    assert plan[:8].tolist() == g["plan_sha_first8"] and plan[-1].tolist() == g["plan_last"]
    assert int((plan.astype(np.int64) * np.arange(1, 7)).sum()) == g["plan_checksum"]
    # product implementation against the same fixture
    r = ZoneRaster(np.broadcast_to(np.zeros((1, 1, 1), np.uint8), (1, H, W)), L, T, RES)
    cfg = {"img_pixels_detection": 512, "margin": m, "output_path": ".", "output_name": "g",
           "reference_modality": "AERIAL_RGBI", "reference_resolution": RES}
    gdf = generate_patches_from_reference(cfg, r, None)
    b = r.bounds
    pplan = tile_plan(gdf, {"left": b.left, "bottom": b.bottom, "right": b.right, "top": b.top}, RES, 512, m)
    assert len(gdf) == g["n_tiles"] and list(gdf["id"][:4]) == g["first_ids"]
    assert int((pplan.astype(np.int64) * np.arange(1, 7)).sum()) == g["plan_checksum"]
    if "plan" in g:
        assert pplan.tolist() == g["plan"]
        got = [[row.left, row.bottom, row.right, row.top] for row in gdf.itertuples()]
        assert got == g["bounds"]                        # bit-exact float64 through JSON round trip


def test_convert_golden():
    g = json.load(open(os.path.join(GOLD, "convert_small.json")))
    logits = np.array(g["logits"], dtype=np.float32)
    assert convert(logits, "argmax").tolist() == g["argmax"]
    assert convert(logits, "class_prob").tolist() == g["class_prob"]
    assert np.array(g["argmax"])[0, 1].tolist() == [min(5, int(np.argmax(logits[:, 1, j]))) for j in range(7)]
    with pytest.raises(ValueError):
        convert(logits, "nope")

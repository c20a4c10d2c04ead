"""Generates tests/golden/*.json from the ORACLE (loop restatement of the reference formulas).
The reference cannot be imported in this image (rasterio/geopandas/smp/timm absent), so these
fixtures pin the oracle's output at the known-answer cases of SURVEY.md H7; the tile counts and
offsets in them were derived independently from slicing.py:51-112 in the survey.
Run:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.convert import convert  # noqa: E402
from oracle.grid import Georef, generate_patches, tile_plan  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
L, T, RES = 700000.0, 6600000.0, 0.2
CASES = [(1000, 700, 64), (2048, 2048, 128), (777, 1300, 40), (10000, 10000, 64), (10000, 10000, 128),
         (10000, 10000, 40), (20000, 20000, 64)]

plans = {}
for (W, H, m) in CASES:
    geo = Georef(L, T, RES, W, H)
    tiles = generate_patches(512, m, RES, geo)
    plan = tile_plan(tiles, geo, 512, m)
    entry = {"n_tiles": len(tiles), "first_ids": [t["id"] for t in tiles[:4]], "last_id": tiles[-1]["id"],
             "plan_sha_first8": plan[:8].tolist(), "plan_last": plan[-1].tolist(),
             "plan_checksum": int((plan.astype(np.int64) * np.arange(1, 7)).sum())}
    if len(tiles) <= 64:
        entry["plan"] = plan.tolist()
        entry["bounds"] = [[t["left"], t["bottom"], t["right"], t["top"]] for t in tiles]
    plans[f"{W}x{H}_m{m}"] = entry
json.dump(plans, open(os.path.join(HERE, "grid_plans.json"), "w"), indent=1)

rng = np.random.default_rng(2025)
logits = (rng.standard_normal((19, 6, 7)) * 3).astype(np.float32)
logits[5, 1, :] = logits.max(axis=0)[1, :]          # ties: first maximal index wins
conv = {"logits": logits.tolist(), "argmax": convert(logits, "argmax").tolist(),
        "class_prob": convert(logits, "class_prob").tolist()}
json.dump(conv, open(os.path.join(HERE, "convert_small.json"), "w"))
print("written", sorted(os.listdir(HERE)))

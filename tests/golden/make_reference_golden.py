"""Generates the golden fixtures of tests/golden/ from the REFERENCE's own code.

    python tests/golden/make_reference_golden.py          (needs /root/reference; run in the build container)

Every value written here comes out of an unmodified function of /root/reference, imported behind the
stand-in modules of tests/reference_stubs.py (rasterio / geopandas / shapely / smp are absent from the image):

  grid_plans.json      slicing.py:20-121 ``generate_patches_from_reference`` on the SURVEY H7 zones, and, per tile, the
                       window the reference's dataset READ (dataset.py:97-115, logged by the in-memory rasterio
                       stand-in) and the window its ``inference_and_write`` WROTE (inference.py:318-352, logged by the
                       recording writer) when the whole zone is pushed through the reference pipeline with a
                       constant-logit model  ->  the integer plan [row0, col0, top_px, left_px, h, w] the kernels consume.
  convert_small.json   postprocess.py:9-30 ``convert`` (argmax with ties, class_prob).
  zone_small.npz       class rasters (argmax) and class_prob planes written by the reference pipeline
                       (reference FLAIR_HUB_Model + load_checkpoint + dataset + inference_and_write) for
                       convnextv2_base-unet on a 1000 x 700 zone, margin 64 -- the product is compared with these on the GPU.
  model_*.npz          see make_model_golden.py (logits of the reference FLAIR_HUB_Model.forward).
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, os.path.dirname(TESTS))
sys.path.insert(0, TESTS)

import reference_stubs as rs  # noqa: E402

L, T, RES = 700000.0, 6600000.0, 0.2
TASK = "AERIAL_LABEL-COSIA"
CASES = [(1000, 700, 64), (2048, 2048, 128), (777, 1300, 40), (10000, 10000, 64), (10000, 10000, 128),
         (10000, 10000, 40), (20000, 20000, 64)]
ZONE_ARCH = "convnextv2_base-unet"      # the headline architecture; 6 tiles of 512 on the CPU
MEANS, STDS = [105.66, 111.35, 102.18, 106.59], [52.23, 45.62, 44.30, 39.78]


class ConstantLogits(torch.nn.Module):
    """Stands in for the network where only the reference's window arithmetic is being recorded."""
    task_nclasses = 19

    def forward(self, inputs):
        x = inputs["AERIAL_RGBI"]
        return {TASK: torch.zeros((x.shape[0], 19, x.shape[-2], x.shape[-1]))}, {}


class LoggingDataset(rs.MemoryDataset):
    """Serves zeros without materialising the raster; logs every read window."""

    def __init__(self, W, H):
        super().__init__(np.broadcast_to(np.zeros((), np.uint8), (4, H, W)), L, T, RES)
        self.reads = []

    def read(self, indexes=None, window=None, out_shape=None, **kw):
        self.reads.append((window.row_off, window.col_off, window.height, window.width))
        return np.zeros(out_shape, np.uint8)


def zonal_cfg(path, W, H, margin, arch="resnet34-unet", weights="w", output_type="argmax"):
    mods = {m: False for m in ("AERIAL_RGBI", "AERIAL-RLT_PAN", "DEM_ELEV", "SPOT_RGBI", "SENTINEL2_TS",
                               "SENTINEL1-ASC_TS", "SENTINEL1-DESC_TS")}
    mods["AERIAL_RGBI"] = True
    return {"output_path": tempfile.gettempdir(), "output_name": f"g{W}x{H}m{margin}", "write_dataframe": False,
            "output_type": output_type, "model_weights": weights, "use_gpu": False, "batch_size": 8, "num_worker": 0,
            "img_pixels_detection": 512, "margin": margin, "output_px_meters": RES, "monotemp_arch": arch,
            "multitemp_model_ref_date": "05-15",
            "modalities": {"inputs": mods, "AERIAL_RGBI": {"input_img_path": path, "channels": [1, 2, 3, 4],
                                                          "normalization": {"type": "custom", "means": MEANS, "stds": STDS}}},
            "tasks": [{"name": TASK, "active": True, "class_names": {i: f"c{i}" for i in range(19)}}]}


def reference_plan(W, H, margin):
    from torch.utils.data import DataLoader
    from flair_zonal_detection import inference as rinf
    from flair_zonal_detection.slicing import generate_patches_from_reference
    path = f"mem://gold_{W}x{H}_{margin}"
    ds = LoggingDataset(W, H)
    rs.register_raster(path, ds)
    cfg = rinf.initialize_geometry_and_resolutions(zonal_cfg(path, W, H, margin))
    cfg["device"] = torch.device("cpu")
    b = ds.bounds
    tiles = generate_patches_from_reference(cfg, path, [rs.Box(b.left - 1, b.bottom - 1, b.right + 1, b.top + 1)])
    dataset = rinf.prep_dataset(cfg, tiles, {"AERIAL_RGBI": 512})
    outs, paths = rinf.init_outputs(cfg, ds, 0)
    rinf.inference_and_write(ConstantLogits(), DataLoader(dataset, batch_size=8, num_workers=0), tiles, cfg, outs, ds)
    w = rs.writer(paths[TASK])
    assert len(ds.reads) == len(w.writes) == len(tiles)
    plan = np.zeros((len(tiles), 6), np.int64)
    for i, ((r, c, hh, ww), (_, r0, c0, h, wd)) in enumerate(zip(ds.reads, w.writes)):
        assert abs(r - round(r)) < 1e-4 and abs(c - round(c)) < 1e-4 and abs(hh - 512) < 1e-4 and abs(ww - 512) < 1e-4
        plan[i] = (round(r), round(c), r0, c0, h, wd)
    return tiles, plan


def main():
    rs.install()
    from flair_zonal_detection.postprocess import convert
    plans = {}
    for (W, H, m) in CASES:
        tiles, plan = reference_plan(W, H, m)
        entry = {"n_tiles": len(tiles), "first_ids": list(tiles["id"][:4]), "last_id": tiles["id"].iloc[-1],
                 "plan_sha_first8": plan[:8].tolist(), "plan_last": plan[-1].tolist(),
                 "plan_checksum": int((plan * np.arange(1, 7)).sum())}
        if len(tiles) <= 64:
            entry["plan"] = plan.tolist()
            entry["bounds"] = [[float(r.left), float(r.bottom), float(r.right), float(r.top)] for r in tiles.itertuples()]
        plans[f"{W}x{H}_m{m}"] = entry
        print(f"{W}x{H} margin {m}: {len(tiles)} tiles")
    plans["_provenance"] = "reference slicing.py / dataset.py / inference.py via tests/golden/make_reference_golden.py"
    json.dump(plans, open(os.path.join(HERE, "grid_plans.json"), "w"), indent=1)

    rng = np.random.default_rng(2025)
    logits = (rng.standard_normal((19, 6, 7)) * 3).astype(np.float32)
    logits[5, 1, :] = logits.max(axis=0)[1, :]          # ties: first maximal index wins
    json.dump({"logits": logits.tolist(), "argmax": convert(logits, "argmax").tolist(),
               "class_prob": convert(logits, "class_prob").tolist(),
               "_provenance": "reference postprocess.py:convert"}, open(os.path.join(HERE, "convert_small.json"), "w"))

    # small zone through the whole reference pipeline
    sys.path.insert(0, TESTS)
    import test_reference_pin as pin
    from flair_for_aigle_b200.synthetic import synthetic_raster
    tmp = tempfile.mkdtemp(prefix="fz_gold_")
    wpath = os.path.join(tmp, "w.safetensors")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_model_golden", os.path.join(HERE, "make_model_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.zone_weights(wpath, ZONE_ARCH, 31)
    arr = synthetic_raster(700, 1000, seed=11)
    out = {}
    for kind in ("argmax", "class_prob"):
        canvas, tiles, _, _ = pin._reference_zone(ZONE_ARCH, arr, 64, kind, wpath, tmp, f"mem://gold_zone_{kind}")
        out[kind] = canvas
        print("zone_small", kind, canvas.shape, "tiles", len(tiles))
    np.savez_compressed(os.path.join(HERE, "zone_small.npz"), argmax=out["argmax"][0],
                        class_prob=out["class_prob"][:, 300:364, 400:528], weights_seed=np.int64(31),
                        raster_seed=np.int64(11))
    print("written", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()

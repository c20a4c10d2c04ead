"""CPU: pins the oracle AND the product's host logic to the reference's OWN Python.

Every test here imports the unmodified modules under /root/reference (behind the stand-in
modules of tests/reference_stubs.py for the third-party packages that are absent from this image)
and compares, on the same inputs,
    reference function  ==  oracle restatement  ==  product drop-in.
Skipped where /root/reference does not exist (the GPU box); the golden fixtures these functions
generate (tests/golden/make_reference_golden.py) are what travels there.
"""
import json
import os
import tempfile

import numpy as np
import pytest
import torch

import reference_stubs as rs

pytestmark = pytest.mark.skipif(not rs.available(), reason="/root/reference not present")

L, T, RES = 700000.0, 6600000.0, 0.2
TASK = "AERIAL_LABEL-COSIA"


@pytest.fixture(scope="module", autouse=True)
def _ref():
    rs.install()


def _cfg(P, margin, res=RES, name="zone"):
    return {"img_pixels_detection": P, "margin": margin, "output_path": tempfile.gettempdir(), "output_name": name,
            "reference_modality": "AERIAL_RGBI", "reference_resolution": res, "write_dataframe": False}


def _ref_tiles(W, H, P, margin, geozone=None, res=RES, left=L, top=T):
    from flair_zonal_detection.slicing import generate_patches_from_reference as ref_gen
    ds = rs.MemoryDataset(np.broadcast_to(np.zeros((), np.uint8), (1, H, W)), left, top, res)
    path = f"mem://pin_{W}x{H}_{left}_{top}_{res}"
    rs.register_raster(path, ds)
    if geozone is None:   # the product script always passes geometries; "whole raster" = a box that contains it
        b = ds.bounds
        geozone = [rs.Box(b.left - 10, b.bottom - 10, b.right + 10, b.top + 10)]
    else:
        geozone = [rs.Box(*geozone)]
    return ref_gen(_cfg(P, margin, res), path, geozone), path


def _product_tiles(W, H, P, margin, geozone=None, res=RES, left=L, top=T):
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    r = ZoneRaster(np.zeros((1, 1, 1), np.uint8), left, top, res)
    r.array = np.broadcast_to(r.array, (1, H, W))
    return generate_patches_from_reference(_cfg(P, margin, res), r, geozone)


COLS = ("left", "bottom", "right", "top", "left_o", "bottom_o", "right_o", "top_o")

def _same_tables(ref, orc, prod):
    """Bit-exact float64 equality of every numeric column, ids and full-tile bounds, in order."""
    assert list(ref["id"]) == [t["id"] for t in orc] == list(prod["id"])
    for c in COLS:
        a = ref[c].to_numpy(np.float64)
        assert np.array_equal(a, np.asarray([t[c] for t in orc], np.float64)), c
        assert np.array_equal(a, prod[c].to_numpy(np.float64)), c
    g = np.asarray([x.bounds for x in ref["geometry"]], np.float64)
    assert np.array_equal(g, np.asarray([t["geometry"] for t in orc], np.float64))
    assert np.array_equal(g, np.asarray([x.bounds for x in prod["geometry"]], np.float64))
    assert list(ref["job_done"]) == list(prod["job_done"]) and list(ref["output_id"]) == list(prod["output_id"])


GRID_CASES = [  # W, H, P, margin, geozone bbox
    (1000, 700, 512, 64, None), (2048, 2048, 512, 128, None), (777, 1300, 512, 40, None),
    (10000, 10000, 512, 64, None), (10000, 10000, 512, 128, None), (10000, 10000, 512, 40, None),
    (700, 512, 512, 0, None), (512, 512, 512, 100, None), (1500, 90, 512, 31, None), (5000, 3000, 512, 17, None),
    (3000, 2500, 256, 20, None), (400, 300, 512, 64, None),
    (4000, 3000, 512, 64, (L + 123.37, T - 501.11, L + 611.93, T - 77.77)),    # geozone inside the raster
    (4000, 3000, 512, 64, (L - 50.0, T - 200.05, L + 300.13, T + 40.0)),       # sticks out top-left
]


@pytest.mark.parametrize("W,H,P,margin,geozone", GRID_CASES)
def test_grid_reference_vs_oracle_vs_product(W, H, P, margin, geozone):
    """slicing.py:20-121 itself, row for row."""
    from oracle.grid import Georef, generate_patches
    ref, path = _ref_tiles(W, H, P, margin, geozone)
    orc = generate_patches(P, margin, RES, Georef(L, T, RES, W, H), geozone, img_path=path, output_name="zone")
    prod = _product_tiles(W, H, P, margin, geozone)
    assert len(ref) == len(orc) == len(prod) and len(ref) > 0
    _same_tables(ref, orc, prod)
    assert list(ref.columns) == list(prod.columns)


def test_grid_random_shapes_reference_vs_oracle_vs_product():
    from oracle.grid import Georef, generate_patches
    rng = np.random.default_rng(7)
    for _ in range(25):
        W, H = int(rng.integers(60, 3000)), int(rng.integers(60, 3000))
        P = int(rng.choice([128, 256, 512]))
        margin = int(rng.integers(0, P // 2 - 1))
        res = float(rng.choice([0.2, 0.5, 1.0, 0.15]))
        left, top = float(rng.uniform(1e5, 9e5)), float(rng.uniform(6e6, 7e6))
        ref, _ = _ref_tiles(W, H, P, margin, None, res, left, top)
        orc = generate_patches(P, margin, res, Georef(left, top, res, W, H), None)
        prod = _product_tiles(W, H, P, margin, None, res, left, top)
        assert len(ref) == len(orc) == len(prod)
        _same_tables(ref, orc, prod)


def test_grid_zone_missing_the_raster_is_empty():
    ref, _ = _ref_tiles(1000, 700, 512, 64, (L - 900.0, T + 100.0, L - 800.0, T + 300.0))
    prod = _product_tiles(1000, 700, 512, 64, (L - 900.0, T + 100.0, L - 800.0, T + 300.0))
    assert len(ref) == 0 and len(prod) == 0


def test_convert_reference_vs_oracle():
    """postprocess.py:9-30."""
    from flair_zonal_detection.postprocess import convert as ref_convert
    from oracle.convert import convert
    rng = np.random.default_rng(3)
    for shape in [(19, 6, 7), (19, 64, 64), (2, 5, 5), (1, 3, 3), (19, 1, 1)]:
        x = (rng.standard_normal(shape) * 4).astype(np.float32)
        if shape[0] > 5:
            x[5, 0, :] = x.max(axis=0)[0, :]       # ties: first maximal index wins
        for kind in ("argmax", "class_prob"):
            a, b = ref_convert(x, kind), convert(x, kind)
            assert a.dtype == b.dtype == np.uint8 and a.shape == b.shape and np.array_equal(a, b)
    for bad in ("softmax", ""):
        with pytest.raises(ValueError, match="Unknown output type"):
            ref_convert(np.zeros((2, 2, 2), np.float32), bad)
        with pytest.raises(ValueError, match="Unknown output type"):
            convert(np.zeros((2, 2, 2), np.float32), bad)
    with pytest.raises(ValueError):
        ref_convert(np.zeros((2, 2), np.float32), "class_prob")


def test_golden_convert_matches_reference():
    from flair_zonal_detection.postprocess import convert as ref_convert
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "convert_small.json")))
    x = np.asarray(g["logits"], np.float32)
    assert np.array_equal(ref_convert(x, "argmax"), np.asarray(g["argmax"], np.uint8))
    assert np.array_equal(ref_convert(x, "class_prob"), np.asarray(g["class_prob"], np.uint8))


def test_norm_reference_vs_oracle_vs_product():
    """norm.py:8-52: 'custom' (float64 per channel), 'scaling' (img_as_float), 'without'."""
    from flair_hub.data.utils_data.norm import norm as ref_norm
    from oracle.pipeline import normalize
    from flair_for_aigle_b200.flair_zonal_detection.dataset import normalization_affine
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, (4, 9, 11), dtype=np.uint8)
    means, stds = [105.66, 111.35, 102.18, 106.59], [52.23, 45.62, 44.30, 39.78]
    a = ref_norm(x.copy(), "custom", means, stds)
    assert a.dtype == np.float64 and np.array_equal(a, normalize(x, means, stds))
    for kind, kw in (("custom", dict(means=means, stds=stds)), ("scaling", {}), ("without", {})):
        want = np.asarray(ref_norm(x.copy(), kind, kw.get("means", []), kw.get("stds", [])), np.float64)
        m, s = normalization_affine({"type": kind, **kw}, 4, x.dtype)
        got = (x.astype(np.float64) - np.asarray(m)[:, None, None]) / np.asarray(s)[:, None, None]
        assert np.allclose(got, want, rtol=0, atol=1e-12), kind
    with pytest.raises(SystemExit):
        ref_norm(x.copy(), "zscore", means, stds)
    with pytest.raises(SystemExit):
        normalization_affine({"type": "zscore"}, 4, x.dtype)
    with pytest.raises(SystemExit):
        normalization_affine({"type": "custom", "means": means, "stds": stds[:3]}, 4, x.dtype)


# ----------------------------------------------------------------------------------------------
# model wiring: the reference's FLAIR_HUB_Model / FusionHandler / FLAIR_Monotemp around the restated smp pieces
# ----------------------------------------------------------------------------------------------
def _zonal_cfg(arch, raster_paths, weights, out_dir, margin=64, P=512, n_cls=19, output_type="argmax", norm=None,
               out_res=None, batch=2):
    """configs/config_model_zonal_segmentation.yaml's schema.  raster_paths: {MOD: (path, channels)}"""
    mods = {m: False for m in ("AERIAL_RGBI", "AERIAL-RLT_PAN", "DEM_ELEV", "SPOT_RGBI", "SENTINEL2_TS",
                               "SENTINEL1-ASC_TS", "SENTINEL1-DESC_TS")}
    cfg = {"output_path": out_dir, "output_name": "pin", "write_dataframe": False, "output_type": output_type,
           "model_weights": weights, "use_gpu": False, "batch_size": batch, "num_worker": 0, "img_pixels_detection": P,
           "margin": margin, "output_px_meters": out_res if out_res is not None else RES, "monotemp_arch": arch,
           "multitemp_model_ref_date": "05-15", "modalities": {"inputs": mods},
           "tasks": [{"name": TASK, "active": True, "class_names": {i: f"c{i}" for i in range(n_cls)}}]}
    for m, (path, chans) in raster_paths.items():
        mods[m] = True
        cfg["modalities"][m] = {"input_img_path": path, "channels": chans,
                                "normalization": norm.get(m) if isinstance(norm, dict) and m in norm else
                                {"type": "custom", "means": [105.66, 111.35, 102.18, 106.59][:len(chans)],
                                 "stds": [52.23, 45.62, 44.30, 39.78][:len(chans)]}}
    if "DEM_ELEV" in raster_paths:
        cfg["modalities"]["DEM_ELEV"].update({"calc_elevation": True, "calc_elevation_stack_dsm": False})
    return cfg


def _model_cfg(arch, mods, n_cls=19):
    """What prepare_model_config produces (model_utils.py:38-109), built by the REFERENCE function."""
    from flair_zonal_detection.model_utils import prepare_model_config
    cfg = _zonal_cfg(arch, {m: ("unused", list(range(1, c + 1))) for m, c in mods.items()}, "w", tempfile.gettempdir(),
                     n_cls=n_cls)
    return prepare_model_config(cfg)


@pytest.mark.parametrize("arch,mods,P", [
    ("resnet34-unet", {"AERIAL_RGBI": 4}, 64),
    ("convnextv2_atto-unet", {"AERIAL_RGBI": 4}, 64),
    ("convnextv2_atto-unet", {"AERIAL_RGBI": 4, "DEM_ELEV": 1}, 64),
    ("swin_tiny_patch4_window7_224-upernet", {"AERIAL_RGBI": 3}, 64),
])
def test_model_wiring_reference_vs_oracle(arch, mods, P):
    """flair_model.py:47-190,357-430,437-547 + monotemp_model.py:34-97 (the reference's classes, unmodified) around
    the restated smp/timm modules == FlairHubOracle: same state_dict keys and shapes, bit-equal logits."""
    from flair_hub.models.flair_model import FLAIR_HUB_Model as RefModel
    from oracle.models import SWIN_CFGS, FlairHubOracle, randomize_
    if arch.split("-")[0] not in SWIN_CFGS and arch.startswith("swin"):
        pytest.skip("oracle restates no small Swin variant")
    torch.manual_seed(0)
    ref = RefModel(_model_cfg(arch, mods), {m: P for m in mods}).eval()
    orc = FlairHubOracle(arch, mods, {TASK: 19}).eval()
    randomize_(orc, seed=5, bf16_exact=False)
    sd_ref, sd_orc = ref.state_dict(), orc.state_dict()
    assert list(sd_ref.keys()) == list(sd_orc.keys())
    assert all(sd_ref[k].shape == sd_orc[k].shape for k in sd_ref)
    ref.load_state_dict(sd_orc, strict=True)
    g = torch.Generator().manual_seed(1)
    batch = {m: torch.randn(2, c, P, P, generator=g) for m, c in mods.items()}
    batch[TASK] = torch.zeros(2, 19, P, P)
    with torch.no_grad():
        (lr, aux_r), (lo, aux_o) = ref(batch), orc(batch)
    assert aux_r == {} and aux_o == {} and list(lr) == list(lo) == [TASK]
    assert lr[TASK].shape == (2, 19, P, P) and torch.equal(lr[TASK], lo[TASK])
    assert ref.task_nclasses == orc.task_nclasses == 19


def test_product_state_dict_layout_matches_reference_model():
    """Key names, order-free, and shapes of the product's parameter tree == the reference module tree's."""
    from flair_hub.models.flair_model import FLAIR_HUB_Model as RefModel
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    for arch, mods in (("resnet34-unet", {"AERIAL_RGBI": 4}), ("convnextv2_base-unet", {"AERIAL_RGBI": 4, "DEM_ELEV": 1}),
                       ("swin_base_patch4_window12_384-upernet", {"AERIAL_RGBI": 4})):
        ref = RefModel(_model_cfg(arch, mods), {m: 512 for m in mods})
        zc = _zonal_cfg(arch, {m: ("unused", list(range(1, c + 1))) for m, c in mods.items()}, "w", tempfile.gettempdir())
        prod = FLAIR_HUB_Model(prepare_model_config(zc), {m: 512 for m in mods})
        a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in prod.state_dict().items()}
        assert a == b, (arch, sorted(set(a) ^ set(b))[:6])


def test_prepare_model_config_reference_vs_product():
    """model_utils.py:38-109 (key-for-key) and compute_patch_sizes (:19-35)."""
    from flair_zonal_detection.model_utils import compute_patch_sizes as ref_sizes, prepare_model_config as ref_prep
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import compute_patch_sizes, prepare_model_config
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster, register_raster
    for res_dem in (0.2, 1.0, 0.4):
        rs.register_raster("mem://pm_rgbi", rs.MemoryDataset(np.zeros((4, 50, 60), np.uint8), L, T, 0.2))
        rs.register_raster("mem://pm_dem", rs.MemoryDataset(np.zeros((1, 10, 12), np.float32), L, T, res_dem))
        register_raster("mem://pm_rgbi", ZoneRaster(np.zeros((4, 50, 60), np.uint8), L, T, 0.2, name="mem://pm_rgbi"))
        register_raster("mem://pm_dem", ZoneRaster(np.zeros((1, 10, 12), np.float32), L, T, res_dem, name="mem://pm_dem"))
        zc = _zonal_cfg("convnextv2_base-unet", {"AERIAL_RGBI": ("mem://pm_rgbi", [1, 2, 3, 4]),
                                                 "DEM_ELEV": ("mem://pm_dem", [1])}, "w.safetensors", tempfile.gettempdir())
        zc["reference_resolution"] = 0.2
        assert ref_prep(zc) == prepare_model_config(zc)
        assert ref_sizes(zc) == compute_patch_sizes(zc)


# ----------------------------------------------------------------------------------------------
# the whole zonal pipeline, reference code end to end
# ----------------------------------------------------------------------------------------------
def _reference_zone(arch, arr, margin, output_type, weights_path, tmp, name, P=512, n_cls=19, out_res=None, batch=2):
    """scripts/run_fast_aigle_segmentation.py:75-119's call order with the reference's own functions."""
    from torch.utils.data import DataLoader
    from flair_zonal_detection import inference as rinf
    from flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
    from flair_zonal_detection.slicing import generate_patches_from_reference
    C, H, W = arr.shape
    ds = rs.MemoryDataset(arr, L, T, RES)
    rs.register_raster(name, ds)
    cfg = _zonal_cfg(arch, {"AERIAL_RGBI": (name, list(range(1, C + 1)))}, weights_path, tmp, margin=margin, P=P,
                     n_cls=n_cls, output_type=output_type, out_res=out_res, batch=batch)
    cfg = rinf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = torch.device("cpu")
    sizes = compute_patch_sizes(cfg)
    model = build_inference_model(cfg, sizes)
    b = ds.bounds
    tiles = generate_patches_from_reference(cfg, name, [rs.Box(b.left - 1, b.bottom - 1, b.right + 1, b.top + 1)])
    dataset = rinf.prep_dataset(cfg, tiles, sizes)
    loader = DataLoader(dataset, batch_size=batch, num_workers=0)
    outs, paths = rinf.init_outputs(cfg, ds, 0)
    rinf.inference_and_write(model, loader, tiles, cfg, outs, ds)
    w = rs.writer(paths[TASK])
    assert w.closed
    return w.canvas, tiles, cfg, model


def _weights(arch, mods, path, seed=3):
    from safetensors.torch import save_file
    from oracle.models import FlairHubOracle, randomize_
    m = FlairHubOracle(arch, mods, {TASK: 19})
    randomize_(m, seed=seed)
    save_file({k: v.contiguous() for k, v in m.state_dict().items()}, path)
    return m.eval()


@pytest.mark.parametrize("output_type,margin,out_res", [("argmax", 16, None), ("class_prob", 16, None),
                                                         ("argmax", 10, 0.4), ("class_prob", 0, 0.5)])
def test_zone_pipeline_reference_vs_oracle(tmp_path, output_type, margin, out_res):
    """Reference: slicing -> MultiModalSlicedDataset (boundless windowed read, normalise) -> DataLoader ->
    FLAIR_HUB_Model -> inference_and_write (crop, convert, zoom, windowed writes, last writer wins) on a 300 x 210
    zone with 128-px tiles == oracle.pipeline.run_zone, byte for byte."""
    from oracle.grid import Georef
    from oracle.pipeline import run_zone
    from flair_for_aigle_b200.synthetic import synthetic_raster
    arch = "resnet34-unet"
    wpath = str(tmp_path / "w.safetensors")
    orc = _weights(arch, {"AERIAL_RGBI": 4}, wpath)
    arr = synthetic_raster(210, 300, seed=4, cell=24)
    ref, tiles, cfg, _ = _reference_zone(arch, arr, margin, output_type, wpath, str(tmp_path), f"mem://zp_{output_type}_{margin}",
                                         P=128, out_res=out_res)
    means, stds = [105.66, 111.35, 102.18, 106.59], [52.23, 45.62, 44.30, 39.78]
    got, _, n = run_zone(orc, arr, Georef(L, T, RES, 300, 210), 128, margin, means, stds, TASK, 19, batch_size=2,
                         output_type=output_type, out_res=out_res)
    assert n == len(tiles) > 4
    got = got[None] if got.ndim == 2 else got
    assert ref.shape == got.shape and ref.dtype == got.dtype == np.uint8
    assert np.array_equal(ref, got)


# ----------------------------------------------------------------------------------------------
# load_checkpoint: reference vs product on crafted checkpoints
# ----------------------------------------------------------------------------------------------
def _both_models(arch, mods, n_cls=19):
    from flair_hub.models.flair_model import FLAIR_HUB_Model as RefModel
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    zc = _zonal_cfg(arch, {m: ("unused", list(range(1, c + 1))) for m, c in mods.items()}, "w", tempfile.gettempdir(),
                    n_cls=n_cls)
    return RefModel(_model_cfg(arch, mods, n_cls), {m: 512 for m in mods}), \
        FLAIR_HUB_Model(prepare_model_config(zc), {m: 512 for m in mods}), _model_cfg(arch, mods, n_cls)


def _load_both(ref, prod, conf, path, seed=11):
    from flair_hub.models.checkpoint import load_checkpoint as ref_load
    from flair_for_aigle_b200.flair_hub.models.checkpoint import load_checkpoint
    conf = dict(conf, paths={"ckpt_model_path": path})
    torch.manual_seed(seed)
    ref_load(conf, ref)
    torch.manual_seed(seed)
    load_checkpoint(conf, prod)
    a, b = ref.state_dict(), prod.state_dict()
    assert set(a) == set(b)
    bad = [k for k in a if not torch.equal(a[k], b[k])]
    assert not bad, bad[:5]
    return a


def _random_state(model, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: (torch.randn(v.shape, generator=g) if v.dtype.is_floating_point else v.clone())
            for k, v in model.state_dict().items()}


def test_load_checkpoint_prefix_head_mismatch_and_torch_load(tmp_path):
    from safetensors.torch import save_file
    ref, prod, conf = _both_models("resnet34-unet", {"AERIAL_RGBI": 4}, n_cls=19)
    sd = _random_state(ref, 1)
    head = f"main_decoders.{TASK}.seg_model.segmentation_head.0"
    # (1) plain safetensors
    p = str(tmp_path / "plain.safetensors")
    save_file(sd, p)
    out = _load_both(ref, prod, conf, p)
    assert torch.equal(out[head + ".weight"], sd[head + ".weight"])
    # (2) Lightning 'model.' prefix + a criterion tensor the module does not have
    p = str(tmp_path / "prefixed.safetensors")
    save_file({**{"model." + k: v for k, v in sd.items()}, f"criterion.{TASK}.weight": torch.ones(19)}, p)
    out = _load_both(ref, prod, conf, p)
    assert torch.equal(out["encoders.AERIAL_RGBI.seg_model.conv1.weight"], sd["encoders.AERIAL_RGBI.seg_model.conv1.weight"])
    # (3) wrong class count in the head (13 classes in the checkpoint): weight Xavier, bias zeros, same RNG draw
    sd13 = dict(sd)
    sd13[head + ".weight"] = torch.randn(13, 16, 3, 3)
    sd13[head + ".bias"] = torch.randn(13)
    p = str(tmp_path / "c13.safetensors")
    save_file(sd13, p)
    out = _load_both(ref, prod, conf, p)
    assert out[head + ".weight"].shape == (19, 16, 3, 3) and float(out[head + ".bias"].abs().max()) == 0.0
    assert prod.last_load_report.reinit_tasks == {TASK} and prod.last_load_report.reinit_tensors == 2
    # (4) head missing altogether, prefixed keys
    p = str(tmp_path / "nohead.safetensors")
    save_file({"model." + k: v for k, v in sd.items() if not k.startswith(head)}, p)
    out = _load_both(ref, prod, conf, p)
    assert float(out[head + ".bias"].abs().max()) == 0.0
    # (5) an encoder tensor of the wrong shape (checkpoint trained on 3 bands) is re-initialised, not loaded
    sd3 = dict(sd)
    sd3["encoders.AERIAL_RGBI.seg_model.conv1.weight"] = torch.randn(64, 3, 7, 7)
    p = str(tmp_path / "c3.safetensors")
    save_file(sd3, p)
    out = _load_both(ref, prod, conf, p)
    assert out["encoders.AERIAL_RGBI.seg_model.conv1.weight"].shape == (64, 4, 7, 7)
    # (6) torch.save'd Lightning checkpoint: {"state_dict": {...}} and a bare dict
    for name, blob in (("lit.ckpt", {"state_dict": {"model." + k: v for k, v in sd.items()}, "epoch": 3}), ("bare.pth", sd)):
        p = str(tmp_path / name)
        torch.save(blob, p)
        out = _load_both(ref, prod, conf, p)
        assert torch.equal(out[head + ".weight"], sd[head + ".weight"])
    # (7) bad path: SystemExit, or a silent return with exit_on_fail=False
    from flair_hub.models.checkpoint import load_checkpoint as ref_load
    from flair_for_aigle_b200.flair_hub.models.checkpoint import load_checkpoint
    bad = dict(conf, paths={"ckpt_model_path": str(tmp_path / "nope.safetensors")})
    for fn, m in ((ref_load, ref), (load_checkpoint, prod)):
        with pytest.raises(SystemExit):
            fn(bad, m)
        assert fn(bad, m, exit_on_fail=False) is None


def test_saved_checkpoint_is_read_by_the_reference_loader(tmp_path):
    """``save_checkpoint`` writes what a Lightning training run of the reference leaves on disk (``model.`` / ``criterion.``
    names, ``state_dict`` inside a ``.ckpt``): the REFERENCE's ``load_checkpoint`` (checkpoint.py:176-290) loads it into the
    reference model and gets the product model's tensors back, for both file types."""
    from flair_for_aigle_b200.flair_hub.models.checkpoint import save_checkpoint
    ref, prod, conf = _both_models("resnet34-unet", {"AERIAL_RGBI": 4}, n_cls=19)
    trained = _random_state(prod, 5)
    prod.load_state_dict(trained, strict=True)
    for name in ("last.ckpt", "last.safetensors"):
        p = save_checkpoint(str(tmp_path / name), prod, {TASK: torch.ones(19)}, epoch=3, global_step=1234,
                            extra={"lr_schedulers": [{"last_epoch": 1234}]})
        fresh_ref, fresh_prod, _ = _both_models("resnet34-unet", {"AERIAL_RGBI": 4}, n_cls=19)
        out = _load_both(fresh_ref, fresh_prod, conf, p)
        assert all(torch.equal(out[k], trained[k]) for k in trained)
    blob = torch.load(str(tmp_path / "last.ckpt"), map_location="cpu", weights_only=False)
    assert blob["epoch"] == 3 and blob["global_step"] == 1234 and blob["lr_schedulers"] == [{"last_epoch": 1234}]
    # the task module's own save (no trainer: weights + the loss weights its criterion carries)
    from flair_for_aigle_b200.flair_hub.tasks.tasks_module import SegmentationTask
    cfg = {"labels": [TASK], "labels_configs": {TASK: {"value_name": list(range(19)), "task_weight": 1.0,
                                                       "value_weights": {"default": 1, "default_exceptions": {18: 0}}}},
           "modalities": {"inputs": {"AERIAL_RGBI": True}, "aux_loss": {"AERIAL_RGBI": False},
                          "modality_dropout": {"AERIAL_RGBI": 0}, "aux_loss_weight": {"AERIAL_RGBI": 1.0}}}
    task = SegmentationTask(prod, cfg)
    p = task.save_checkpoint(str(tmp_path / "task.ckpt"), epoch=1)
    tb = torch.load(p, map_location="cpu", weights_only=False)
    w = tb["state_dict"][f"criterion.{TASK}.weight"]
    assert w.shape == (19,) and float(w[18]) == 0.0 and float(w[:18].min()) == 1.0 and "optimizer_states" not in tb
    fresh_ref, fresh_prod, _ = _both_models("resnet34-unet", {"AERIAL_RGBI": 4}, n_cls=19)
    out = _load_both(fresh_ref, fresh_prod, conf, p)
    assert all(torch.equal(out[k], trained[k]) for k in trained)
    assert f"criterion.{TASK}.weight" in blob["state_dict"] and all(k.startswith(("model.", "criterion.")) for k in blob["state_dict"])


def test_load_checkpoint_swin_bias_table_resize(tmp_path):
    """checkpoint.py:33-56,265-271: a window-7 checkpoint's (13*13, heads) tables resized bicubically to the
    window-12 model's (23*23, heads); a non-square table falls back to re-initialisation ('bias' in the name -> zeros)."""
    from safetensors.torch import save_file
    arch = "swin_base_patch4_window12_384-upernet"
    ref, prod, conf = _both_models(arch, {"AERIAL_RGBI": 4})
    sd = _random_state(ref, 2)
    tables = [k for k in sd if k.endswith("relative_position_bias_table")]
    assert len(tables) == 24 and all(sd[k].shape[0] == 23 * 23 for k in tables)
    small = dict(sd)
    for k in tables[:-1]:
        small[k] = torch.randn(13 * 13, sd[k].shape[1])
    small[tables[-1]] = torch.randn(150, sd[tables[-1]].shape[1])          # not a square
    p = str(tmp_path / "w7.safetensors")
    save_file(small, p)
    out = _load_both(ref, prod, conf, p)
    k = tables[0]
    want = torch.nn.functional.interpolate(small[k].reshape(1, 13, 13, -1).permute(0, 3, 1, 2), size=(23, 23),
                                           mode="bicubic", align_corners=False).permute(0, 2, 3, 1).reshape(529, -1)
    assert torch.equal(out[k], want)
    assert float(out[tables[-1]].abs().max()) == 0.0
    assert len(prod.last_load_report.resized) == 23 and prod.last_load_report.reinit_tensors == 1


# ----------------------------------------------------------------------------------------------
# training step: loss side
# ----------------------------------------------------------------------------------------------
def test_training_step_loss_reference_vs_oracle():
    """tasks_module.py:133-167 (SegmentationTask.step) + module_setup.py:119-200 (FLAIRLosses), the reference's
    classes, on the oracle model == oracle.training.step; default class weights == the product's FLAIRLosses."""
    from flair_hub.tasks.module_setup import FLAIRLosses as RefLosses
    from flair_hub.tasks.tasks_module import SegmentationTask
    from oracle.models import FlairHubOracle, randomize_
    from oracle.training import default_class_weights, init_optimizer, step as oracle_step
    from flair_for_aigle_b200.flair_hub.tasks.module_setup import FLAIRLosses
    mods = {"AERIAL_RGBI": 4, "DEM_ELEV": 1}
    cfg = {"labels": [TASK],
           "labels_configs": {TASK: {"value_name": list(range(19)), "task_weight": 0.7,
                                     "value_weights": {"default": 1, "default_exceptions": {15: 0, 16: 0, 17: 0, 18: 0.5}}}},
           "modalities": {"inputs": {m: True for m in mods}, "aux_loss": {m: False for m in mods},
                          "modality_dropout": {m: 0 for m in mods}, "aux_loss_weight": {m: 1.0 for m in mods}},
           "hyperparams": {"optimizer": "adamw", "learning_rate": 5e-5, "optim_weight_decay": 0.01, "optim_betas": [0.9, 0.999]}}
    model = FlairHubOracle("convnextv2_atto-unet", mods, {TASK: 19})
    randomize_(model, seed=9, bf16_exact=False)
    model.train()
    real_forward = model.forward
    model.forward = lambda batch, apply_mod_dropout=False: real_forward(batch)
    ref_losses = RefLosses(cfg)
    task = SegmentationTask(model, cfg, criterion=ref_losses.get_losses())
    w_ref = ref_losses.get_default_weights(TASK)
    assert torch.equal(w_ref, default_class_weights(cfg["labels_configs"][TASK]))
    assert torch.equal(w_ref, FLAIRLosses(cfg).default_weights[TASK])
    g = torch.Generator().manual_seed(2)
    batch = {m: torch.randn(2, c, 64, 64, generator=g) for m, c in mods.items()}
    batch[TASK] = torch.nn.functional.one_hot(torch.randint(0, 19, (2, 64, 64), generator=g), 19).permute(0, 3, 1, 2).float()
    loss_r, preds_r, targets_r = task.step(batch, training=True)
    loss_o, preds_o, targets_o = oracle_step(model, batch, cfg)
    assert torch.equal(loss_r, loss_o) and torch.equal(preds_r[TASK], preds_o[TASK])
    assert torch.equal(targets_r[TASK], targets_o[TASK]) and targets_r[TASK].dtype == torch.int32
    # optimizer: tasks_module.py:377-391
    opt_r = task._init_optimizer(cfg["hyperparams"])
    opt_o = init_optimizer(cfg["hyperparams"], model.parameters())
    assert type(opt_r) is type(opt_o) is torch.optim.AdamW
    assert {k: v for k, v in opt_r.defaults.items()} == {k: v for k, v in opt_o.defaults.items()}


def test_reference_aux_loss_is_identically_zero():
    """Why the training engine builds no auxiliary decoders: in the reference as it runs, ``_compute_aux_loss``
    (tasks_module.py:169-194) looks for the TASK name among the keys of ``dict_logits_aux``, whose keys are
    ``aux_<mod>_<task>`` (flair_model.py:384,402) -- the branch is never taken, the auxiliary logits never reach the loss and
    the auxiliary decoders never receive a gradient.  Pinned here on the reference's own method with auxiliary losses switched
    ON in the config and wildly wrong auxiliary logits."""
    from flair_hub.tasks.module_setup import FLAIRLosses as RefLosses
    from flair_hub.tasks.tasks_module import SegmentationTask
    mods = {"AERIAL_RGBI": 4, "DEM_ELEV": 1}
    cfg = {"labels": [TASK],
           "labels_configs": {TASK: {"value_name": list(range(19)), "task_weight": 1.0,
                                     "value_weights": {"default": 1, "default_exceptions": {}}}},
           "modalities": {"inputs": {m: True for m in mods}, "aux_loss": {m: True for m in mods},
                          "modality_dropout": {m: 0 for m in mods}, "aux_loss_weight": {m: 1.0 for m in mods}},
           "hyperparams": {"optimizer": "adamw", "learning_rate": 5e-5, "optim_weight_decay": 0.01, "optim_betas": [0.9, 0.999]}}
    task = SegmentationTask(torch.nn.Identity(), cfg, criterion=RefLosses(cfg).get_losses())
    assert task.aux_loss_modalities == list(mods)                        # the auxiliary losses ARE configured
    g = torch.Generator().manual_seed(4)
    targets = torch.randint(0, 19, (2, 16, 16), generator=g)
    aux = {f"aux_{m}_{TASK}": torch.randn(2, 19, 16, 16, generator=g) * 50 for m in mods}
    out = task._compute_aux_loss(aux, TASK, targets)
    assert float(out) == 0.0 and out.dtype == torch.float32


def test_modality_dropout_draws_reference_vs_oracle_vs_product():
    """flair_model.py:406-408 + :330-354 on the reference's own method: with the same seeds the oracle's restatement and the
    product's ``draw_modality_dropout`` take the same decisions and produce the same noise (torch's generators on both
    sides), over enough seeds that 'both kept', 'one dropped' and 'both dropped' all occur."""
    import random
    from flair_hub.models.flair_model import FLAIR_HUB_Model as RefModel
    from oracle.models import FlairHubOracle
    from flair_for_aigle_b200.flair_hub.models.flair_model import draw_modality_dropout
    B, P, dims = 2, 64, (40, 80, 160, 320)
    shapes = {"AERIAL_RGBI": [(B, 4, P, P), (B, 0, P // 2, P // 2)] + [(B, c, P // (4 << i), P // (4 << i)) for i, c in enumerate(dims)],
              "DEM_ELEV": [(B, 1, P, P), (B, 0, P // 2, P // 2)] + [(B, c, P // (4 << i), P // (4 << i)) for i, c in enumerate(dims)]}
    seen = set()
    for seed in range(12):
        outs = []
        for which in ("reference", "oracle", "product"):
            random.seed(seed)
            torch.manual_seed(seed)
            fmaps = {k: [torch.zeros(s) for s in v] for k, v in shapes.items()}
            if which == "product":
                dropped = draw_modality_dropout(shapes, "cpu")
            else:
                probs = {key: random.uniform(0, 1) for key in fmaps.keys()}               # flair_model.py:407
                fn = RefModel.modality_dropout if which == "reference" else FlairHubOracle.modality_dropout
                res = fn(None, fmaps, probs) if which == "reference" else fn(fmaps, probs)
                dropped = {k: v for k, v in res.items() if any(t.abs().sum() > 0 for t in v)}
            outs.append({k: [t.detach() for t in v] for k, v in dropped.items()})
        ref, ora, prod = outs
        assert set(ref) == set(ora) == set(prod)
        seen.add(len(ref))
        for k in ref:
            assert all(torch.equal(a, b) and torch.equal(a, c) for a, b, c in zip(ref[k], ora[k], prod[k]))
            bound = (6.0 / ((B + dims[0]) * (P // 4) ** 2)) ** 0.5                         # xavier: fan_in + fan_out = (C + B) h w
            assert float(ref[k][2].abs().max()) <= bound and ref[k][1].numel() == 0
    assert seen == {0, 1, 2}

"""Zonal pipeline with two mono-temporal modalities: AERIAL_RGBI (uint8) + DEM_ELEV (float32 raster on the
same grid) -> per-modality windows (dataset.py:89-124,174-209) -> two encoders + FusionHandler + U-Net -> class raster,
against the oracle pipeline driven with both modalities."""
import numpy as np
import pytest
import torch
from parity import (CLASS_AGREEMENT, CLASS_AGREEMENT_CONFIDENT, CLASS_AGREEMENT_FUSED, LOGIT_MAX_ABS,  # noqa: E402,F401
                    LOGIT_MEAN_ABS)

L, T, RES = 700000.0, 6600000.0, 0.2
TASK = "AERIAL_LABEL-COSIA"
DEM_MEAN, DEM_STD = [12.5], [7.0]


def _config(wpath, out_dir, aerial, dem, batch=2):
    import bench
    c = bench.zonal_config(wpath, out_dir, aerial, batch)
    c["modalities"]["inputs"]["DEM_ELEV"] = True
    c["modalities"]["DEM_ELEV"] = {"input_img_path": dem, "channels": [1], "calc_elevation": True,
                                   "calc_elevation_stack_dsm": False,
                                   "normalization": {"type": "custom", "means": DEM_MEAN, "stds": DEM_STD}}
    return c


def _rasters(tag, W=1000, H=700, dem_res=RES, dem_shift=0):
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster, register_raster
    from flair_for_aigle_b200.synthetic import synthetic_raster
    arr = synthetic_raster(H, W, seed=11)
    rng = np.random.default_rng(4)
    k = RES / dem_res
    dh, dw = int(round(H * k)), int(round(W * k))
    yy, xx = np.mgrid[:dh, :dw]
    dem = (12.0 + 6.0 * np.sin(xx / 37.0) * np.cos(yy / 53.0) + rng.standard_normal((dh, dw))).astype(np.float32)[None]
    a, d = f"mem://mm_aerial_{tag}", f"mem://mm_dem_{tag}"
    register_raster(a, ZoneRaster(arr, L, T, RES, name=a))
    register_raster(d, ZoneRaster(dem, L + dem_shift, T, dem_res, name=d))
    return arr, dem, a, d


def test_modality_read_plans(tmp_path):
    """Host logic (dataset.py:97 ``from_bounds`` per modality): a second raster on the same grid reads the reference windows
    as whole-pixel copies; a raster shifted by half a pixel, or at another pixel size, gets FRACTIONAL windows -> the
    resampled read (rasterio bilinear, fz_gather_tiles_resampled); the windows equal rasterio's float arithmetic."""
    import reference_stubs as rs
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.raster import open_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    arr, dem, a, d = _rasters("cpu")
    cfg = inf.initialize_geometry_and_resolutions(_config("unused", str(tmp_path), a, d))
    tiles = generate_patches_from_reference(cfg, a, None)
    ds = inf.prep_dataset(cfg, tiles, {"AERIAL_RGBI": 512, "DEM_ELEV": 512})
    assert np.array_equal(ds.modality_origins("DEM_ELEV"), ds.plan()[:, :2])
    assert ds.modality_origins("AERIAL_RGBI").dtype == np.int32
    assert ds.modality_read_plan("DEM_ELEV")[0] == "aligned"
    _, _, a2, d2 = _rasters("shift", dem_shift=0.1)                   # half a pixel off
    cfg2 = dict(cfg)
    cfg2["modalities"] = {**cfg["modalities"], "DEM_ELEV": {**cfg["modalities"]["DEM_ELEV"], "input_img_path": d2}}
    ds2 = inf.prep_dataset(cfg2, tiles, {"AERIAL_RGBI": 512, "DEM_ELEV": 512})
    kind, win = ds2.modality_read_plan("DEM_ELEV")
    assert kind == "resampled" and win.dtype == np.float64 and win.shape == (len(tiles), 4)
    assert np.allclose(win[:, 1] - ds2.plan()[:, 1], -0.5, atol=1e-6) and np.allclose(win[:, 2:], 512.0, atol=1e-6)
    with pytest.raises(ValueError, match="fractional"):
        ds2.modality_origins("DEM_ELEV")
    # a DEM at 1 m under the 0.2 m ortho: window 102.4 px -> patch 102 (model_utils.py:19-35); windows == rasterio's
    _, _, a3, d3 = _rasters("coarse", dem_res=1.0)
    cfg3 = dict(cfg)
    cfg3["modalities"] = {**cfg["modalities"], "DEM_ELEV": {**cfg["modalities"]["DEM_ELEV"], "input_img_path": d3}}
    ds3 = inf.prep_dataset(cfg3, tiles, {"AERIAL_RGBI": 512, "DEM_ELEV": 102})
    kind, win = ds3.modality_read_plan("DEM_ELEV")
    assert kind == "resampled" and np.allclose(win[:, 2:], 102.4, atol=1e-9)
    r = open_raster(d3)
    t = rs.from_origin(r.left, r.top, 1.0, 1.0)
    for i, g in enumerate(tiles["geometry"]):
        w = rs.window_from_bounds(*g.bounds, transform=t)
        assert (w.row_off, w.col_off, w.height, w.width) == tuple(win[i])


@pytest.mark.gpu
def test_two_modality_zone_vs_oracle(cuda, tmp_path):
    from safetensors.torch import save_file
    from oracle.convert import write_tiles
    from oracle.grid import Georef, generate_patches, tile_plan
    from oracle.models import FlairHubOracle
    from oracle.pipeline import load_batch
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import (build_inference_model, compute_patch_sizes,
                                                                        prepare_model_config)
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS, randomize_state_
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    arr, dem, a, d = _rasters("gpu")
    wpath = str(tmp_path / "w2.safetensors")
    cfg = _config(wpath, str(tmp_path), a, d)
    proto = FLAIR_HUB_Model(prepare_model_config(cfg), {"AERIAL_RGBI": 512, "DEM_ELEV": 512}, max_batch=2)
    sd = proto.state_dict()
    randomize_state_(sd, seed=21)
    save_file({k: v.contiguous() for k, v in sd.items()}, wpath)
    cfg = inf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = cuda
    sizes = compute_patch_sizes(cfg)
    assert sizes == {"AERIAL_RGBI": 512, "DEM_ELEV": 512}
    model = build_inference_model(cfg, sizes).to(cuda)
    assert list(model.active_mono) == ["AERIAL_RGBI", "DEM_ELEV"]
    tiles = generate_patches_from_reference(cfg, a, None)
    ds = inf.prep_dataset(cfg, tiles, sizes)
    RasterSink.write_files = False
    outs, _ = inf.init_outputs(cfg, a, 0)
    inf.inference_and_write(model, ds, tiles, cfg, outs, a)
    got = outs[TASK].to_host()[0]

    oracle = FlairHubOracle("convnextv2_base-unet", {"AERIAL_RGBI": 4, "DEM_ELEV": 1}, {TASK: 19}).eval()
    oracle.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    oracle = oracle.to(cuda)
    geo = Georef(L, T, RES, 1000, 700)
    plan = tile_plan(generate_patches(512, 64, RES, geo), geo, 512, 64, None)
    want = np.zeros((700, 1000), np.uint8)
    with torch.no_grad():
        for s in range(0, plan.shape[0], 2):
            idx = list(range(s, min(s + 2, plan.shape[0])))
            batch = load_batch(arr, plan, idx, 512, DEFAULT_MEANS, DEFAULT_STDS, TASK, 19)
            batch["DEM_ELEV"] = load_batch(dem, plan, idx, 512, DEM_MEAN, DEM_STD, TASK, 19, mod="DEM_ELEV")["DEM_ELEV"]
            logits, _ = oracle({k: v.to(cuda) for k, v in batch.items()})
            write_tiles(logits[TASK].cpu().numpy(), plan[idx], 64, want, "argmax")
    agree = (got == want).mean()
    print(f"two-modality zone: class agreement with the oracle pipeline {agree:.5f}")
    assert agree >= CLASS_AGREEMENT_FUSED          # two encoders feed the fusion: measured 0.99817 (fp16 operands)
    # the DEM really reaches the prediction: a different elevation raster changes the class map
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster, register_raster
    register_raster(d, ZoneRaster(dem[:, ::-1].copy() * 3.0, L, T, RES, name=d))
    ds2 = inf.prep_dataset(cfg, tiles, sizes)
    outs2, _ = inf.init_outputs(cfg, a, 0)
    inf.inference_and_write(model, ds2, tiles, cfg, outs2, a)
    assert (outs2[TASK].to_host()[0] != got).mean() > 0.01


def _resample_case(rng, dtype, H, W, n, ps, ratio_lo, ratio_hi):
    C = 3 if dtype == np.uint8 else 1
    raster = rng.integers(0, 256, (C, H, W), dtype=np.uint8) if dtype == np.uint8 else \
        (rng.standard_normal((C, H, W)) * 20 + 100).astype(np.float32)
    win = np.zeros((n, 4))
    for i in range(n):
        r = rng.uniform(ratio_lo, ratio_hi)
        h = w = ps * r
        win[i] = (rng.uniform(-0.4 * h, H - 0.6 * h), rng.uniform(-0.4 * w, W - 0.6 * w), h, w)   # sticks out of the raster
    return raster, win


def test_oracle_resampled_read_basics():
    """oracle/resample.py: an integer-aligned 1:1 window is a plain copy with zero fill; a half-pixel shift is the mean of two
    neighbours; a 2:1 downsample of a constant is the constant; fill pixels take part like data."""
    from oracle.grid import read_tile
    from oracle.resample import read_resampled
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, (2, 40, 50), dtype=np.uint8)
    assert np.array_equal(read_resampled(src, -5, 30, 32, 32, 32, 32), read_tile(src, -5, 30, 32))
    f = (rng.standard_normal((1, 20, 20)) * 3).astype(np.float32)
    half = read_resampled(f, 2.0, 3.5, 8, 8, 8, 8)
    assert np.allclose(half[0], 0.5 * (f[0, 2:10, 3:11].astype(np.float64) + f[0, 2:10, 4:12]), atol=1e-6)
    const = np.full((1, 64, 64), 7.0, np.float32)
    assert np.allclose(read_resampled(const, 8.0, 8.0, 32.0, 32.0, 16, 16), 7.0)
    edge = read_resampled(const, -4.25, 8.0, 16.0, 16.0, 16, 16)         # rows above the raster are fill (0)
    assert np.allclose(edge[0, :4], 0.0) and np.allclose(edge[0, 4], 0.75 * 7.0) and np.allclose(edge[0, 5:], 7.0)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,ps,lo,hi", [(np.float32, 102, 1.0039, 1.0040), (np.uint8, 128, 0.4, 0.9), (np.float32, 64, 1.5, 4.7),
                                            (np.uint8, 96, 1.0, 1.0), (np.float32, 40, 7.0, 9.4)])
def test_resampled_gather_kernel_vs_oracle(cuda, dtype, ps, lo, hi):
    """fz_gather_tiles_resampled == oracle.resample.read_resampled + norm.py's float64 (x - mean) / std, to 1 float32 ulp of
    the normalised value (same operations in double; GDAL's half-up rounding for uint8 rasters)."""
    from oracle.pipeline import normalize
    from oracle.resample import read_resampled
    from flair_for_aigle_b200 import native as nv
    rng = np.random.default_rng(ps)
    raster, win = _resample_case(rng, dtype, 300, 340, 6, ps, lo, hi)
    if lo == hi == 1.0:
        win[:, :2] = np.round(win[:, :2]) + 0.25                    # 1:1 scale but a quarter-pixel shift
    C = raster.shape[0]
    means = [100.5, 90.25, 110.0][:C]
    stds = [50.0, 45.5, 40.25][:C]
    mean = torch.tensor(means, dtype=torch.float32, device=cuda)
    std = torch.tensor(stds, dtype=torch.float32, device=cuda)
    got = nv.gather_tiles_resampled(torch.from_numpy(raster).to(cuda), torch.from_numpy(win).to(cuda), ps, mean, std)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    for i in range(win.shape[0]):
        patch = read_resampled(raster, win[i, 0], win[i, 1], win[i, 2], win[i, 3], ps, ps)
        want = normalize(patch, means, stds).astype(np.float32)
        d = np.abs(got[i] - want)
        tol = np.spacing(np.abs(want).max().astype(np.float32)) * (1 if dtype == np.float32 else 0)
        if dtype == np.uint8:
            # a resampled value within 1e-9 of x.5 may round the other way: at most a handful of pixels differ by one level
            assert (d > 1e-6).mean() < 1e-3 and d.max() <= 1.0 / min(stds) + 1e-6
        else:
            assert d.max() <= 2 * tol + 1e-6 * np.abs(want).max(), (i, d.max())


@pytest.mark.gpu
def test_zone_batches_with_a_coarser_dem_equal_the_reference_style_read(cuda, tmp_path):
    """_iter_batches (the generic path's feeder) for AERIAL_RGBI at 0.2 m + DEM_ELEV at 1 m: the DEM tiles are the resampled
    102 x 102 reads the reference's dataset produces (dataset.py:89-124 through the rasterio stand-in of tests/reference_stubs.py
    = oracle.resample), normalised."""
    import reference_stubs as rs
    from oracle.pipeline import normalize
    from oracle.resample import read_resampled
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.raster import open_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    arr, dem, a, d = _rasters("coarse_gpu", dem_res=1.0)
    cfg = inf.initialize_geometry_and_resolutions(_config("unused", str(tmp_path), a, d, batch=3))
    cfg["device"] = cuda
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import compute_patch_sizes
    sizes = compute_patch_sizes(cfg)
    assert sizes == {"AERIAL_RGBI": 512, "DEM_ELEV": 102}
    tiles = generate_patches_from_reference(cfg, a, None)
    ds = inf.prep_dataset(cfg, tiles, sizes)

    class _Mods:
        active_mono = ["AERIAL_RGBI", "DEM_ELEV"]
    batches = list(inf._iter_batches(None, ds, _Mods(), cfg, cuda))
    assert sum(b["index"].numel() for b in batches) == len(tiles)
    r = open_raster(d)
    t = rs.from_origin(r.left, r.top, 1.0, 1.0)
    k = 0
    for b in batches:
        assert tuple(b["DEM_ELEV"].shape[1:]) == (1, 102, 102) and tuple(b["AERIAL_RGBI"].shape[1:]) == (4, 512, 512)
        x = b["DEM_ELEV"].cpu().numpy()
        for j in range(x.shape[0]):
            w = rs.window_from_bounds(*tiles["geometry"].iloc[k].bounds, transform=t)
            want = normalize(read_resampled(dem, w.row_off, w.col_off, w.height, w.width, 102, 102), DEM_MEAN, DEM_STD)
            assert np.abs(x[j] - want.astype(np.float32)).max() <= 2e-6 * np.abs(want).max()
            k += 1

"""output_px_meters != reference resolution (inference.py:165-194, 212-226, 299-352): host pieces on CPU, the zonal
path on the GPU against the oracle pipeline (which calls scipy.ndimage.zoom like the reference)."""
import numpy as np
import pytest
import torch
from parity import (CLASS_AGREEMENT, CLASS_AGREEMENT_CONFIDENT, CLASS_AGREEMENT_FUSED, LOGIT_MAX_ABS,  # noqa: E402,F401
                    LOGIT_MEAN_ABS)

L, T, RES = 700000.0, 6600000.0, 0.2
TASK = "AERIAL_LABEL-COSIA"


@pytest.mark.parametrize("size,scale", [(384, 0.5), (384, 0.4), (384, 2.0), (384, 0.8), (256, 1.25), (100, 0.333)])
def test_zoom_map_is_scipy_zoom(size, scale):
    from scipy.ndimage import zoom
    from flair_for_aigle_b200.flair_zonal_detection.inference import resample_prediction
    from flair_for_aigle_b200.flair_zonal_detection.slicing import zoom_map
    zm = zoom_map(size, scale)
    assert len(zm) == int(round(size * scale)) and zm.min() >= -1 and zm.max() <= size - 1
    rng = np.random.default_rng(size)
    lab = rng.integers(0, 19, (size, size), dtype=np.uint8)
    assert np.array_equal(resample_prediction(lab, scale), zoom(lab, zoom=scale, order=0))
    lg = rng.standard_normal((5, size, size)).astype(np.float32)
    assert np.array_equal(resample_prediction(lg, scale), zoom(lg, zoom=(1, scale, scale), order=0))


@pytest.mark.parametrize("out_res", [0.4, 0.5, 0.25])
def test_rescaled_tile_plan_matches_oracle(out_res):
    from oracle.grid import Georef, generate_patches, tile_plan as oracle_plan
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster, register_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference, tile_plan
    W, H = 2300, 1700
    geo = Georef(L, T, RES, W, H)
    want = oracle_plan(generate_patches(512, 64, RES, geo), geo, 512, 64, out_res)
    name = f"mem://plan_{out_res}"
    register_raster(name, ZoneRaster(np.broadcast_to(np.zeros((1, 1, 1), np.uint8), (4, H, W)), L, T, RES))
    import bench
    cfg = bench.zonal_config("unused", "/tmp", name, 4)
    cfg["reference_resolution"] = RES
    left, bottom, right, top = geo.bounds
    ib = {"left": left, "bottom": bottom, "right": right, "top": top}
    tiles = generate_patches_from_reference(cfg, name, None)
    got = tile_plan(tiles, ib, RES, 512, 64, out_res)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("out_res,output_type", [(0.4, "argmax"), (0.5, "argmax"), (0.25, "argmax"), (0.4, "class_prob")])
def test_zone_with_rescaled_output(cuda, tmp_path, out_res, output_type):
    import bench
    from safetensors.torch import load_file
    from oracle.grid import Georef
    from oracle.models import FlairHubOracle
    from oracle.pipeline import run_zone
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink, ZoneRaster, register_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS, synthetic_raster
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    wpath = str(tmp_path / "w.safetensors")
    bench.make_weights(wpath, seed=7)
    arr = synthetic_raster(700, 1000, seed=11)
    name = f"mem://rescale_{out_res}_{output_type}"
    register_raster(name, ZoneRaster(arr, L, T, RES, name=name))
    cfg = bench.zonal_config(wpath, str(tmp_path), name, 4)
    cfg["output_px_meters"] = out_res
    cfg["output_type"] = output_type
    cfg = inf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = cuda
    sizes = compute_patch_sizes(cfg)
    model = build_inference_model(cfg, sizes).to(cuda)
    tiles = generate_patches_from_reference(cfg, name, None)
    ds = inf.prep_dataset(cfg, tiles, sizes)
    RasterSink.write_files = False
    outs, _ = inf.init_outputs(cfg, name, 0)
    inf.inference_and_write(model, ds, tiles, cfg, outs, name)
    got = outs[TASK].to_host()
    oracle = FlairHubOracle("convnextv2_base-unet", {"AERIAL_RGBI": 4}, {TASK: 19}).eval()
    oracle.load_state_dict(load_file(wpath), strict=True)
    ref, _, _ = run_zone(oracle.to(cuda), arr, Georef(L, T, RES, 1000, 700), 512, 64, DEFAULT_MEANS, DEFAULT_STDS, TASK,
                         19, batch_size=2, output_type=output_type, device="cuda", out_res=out_res)
    if output_type == "argmax":
        got = got[0]
        assert got.shape == ref.shape == (int(round(700 * RES / out_res)), int(round(1000 * RES / out_res)))
        agree = (got == ref).mean()
        print(f"out_res {out_res}: class agreement with the oracle pipeline {agree:.5f}")
        assert agree >= CLASS_AGREEMENT
    else:
        assert got.shape == ref.shape
        d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
        print(f"class_prob at out_res {out_res}: mean |d| {d.mean():.3f} LSB, max {d.max()}")
        assert d.mean() < 1.0       # bf16 logits noise through softmax*255


@pytest.mark.gpu
@pytest.mark.parametrize("out_res", [0.4, 0.5, 0.25, 0.3])
def test_zoom_write_is_bit_exact_on_identical_logits(cuda, out_res):
    """Same logits in, same bytes out: fz_crop_zoom_write vs the oracle's scipy-based write (incl. the tiles clipped at the
    raster edge, the last-writer rule in output pixels and scipy's constant-fill edge pixels)."""
    from oracle.convert import write_tiles_rescaled
    from oracle.grid import Georef, generate_patches, tile_plan
    from flair_for_aigle_b200 import native as nv
    from flair_for_aigle_b200.flair_zonal_detection.slicing import ownership_windows, zoom_map
    W, H, P, m = 1000, 700, 512, 64
    geo = Georef(L, T, RES, W, H)
    plan = tile_plan(generate_patches(P, m, RES, geo), geo, P, m, out_res)
    n = plan.shape[0]
    rng = np.random.default_rng(int(out_res * 100))
    logits = rng.standard_normal((n, 19, P, P)).astype(np.float32)
    logits[:, 7, ::5] = logits.max(axis=1)[:, ::5]                 # ties: first maximal index wins
    oh, ow = int(round(H * RES / out_res)), int(round(W * RES / out_res))
    want = np.full((oh, ow), 255, np.uint8)
    write_tiles_rescaled(logits, plan, m, want, "argmax", RES / out_res)
    got = torch.full((oh, ow), 255, dtype=torch.uint8, device=cuda)
    zm = torch.from_numpy(zoom_map(P - 2 * m, RES / out_res)).to(cuda)
    nv.crop_zoom_write(0, torch.from_numpy(logits).to(cuda), nv.NCHW, m, torch.from_numpy(plan).to(cuda),
                       torch.from_numpy(ownership_windows(plan)).to(cuda), zm, got)
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("out_res", [0.4, 0.5, 0.25])
def test_zoom_accumulate_matches_oracle_on_identical_logits(cuda, out_res):
    """fz_crop_zoom_accumulate vs the oracle's scipy zoom + softmax accumulation on the same logits: fp32, 1e-6 absolute
    per contribution (expf ulp), clamped edge tiles overlapping their neighbours included."""
    from oracle.convert import blend_accumulate_rescaled
    from oracle.grid import Georef, generate_patches, tile_plan
    from flair_for_aigle_b200 import native as nv
    from flair_for_aigle_b200.flair_zonal_detection.slicing import zoom_map
    W, H, P, m = 1000, 700, 512, 64
    geo = Georef(L, T, RES, W, H)
    plan = tile_plan(generate_patches(P, m, RES, geo), geo, P, m, out_res)
    rng = np.random.default_rng(int(out_res * 1000))
    logits = (3 * rng.standard_normal((plan.shape[0], 19, P, P))).astype(np.float32)
    oh, ow = int(round(H * RES / out_res)), int(round(W * RES / out_res))
    want = np.zeros((19, oh, ow), np.float32)
    blend_accumulate_rescaled(logits, plan, m, want, RES / out_res)
    got = torch.zeros((19, oh, ow), dtype=torch.float32, device=cuda)
    zm = torch.from_numpy(zoom_map(P - 2 * m, RES / out_res)).to(cuda)
    nv.crop_zoom_accumulate(torch.from_numpy(logits).to(cuda), nv.NCHW, m, torch.from_numpy(plan).to(cuda), zm, got)
    got = got.cpu().numpy()
    cover = want.sum(axis=0)
    assert cover.min() > 0.99 and np.allclose(cover, np.round(cover), atol=1e-4)       # every pixel covered 1..4 times
    assert np.abs(got - want).max() < 4e-6


@pytest.mark.gpu
def test_inference_accumulating_variant_on_rescaled_grid(cuda, tmp_path):
    """inference() (inference.py:468-564) at output_px_meters = 0.4: canvas on the out_res grid, its labels agree with
    inference_and_write's raster wherever exactly one tile covers the pixel."""
    import bench
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink, ZoneRaster, register_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    from flair_for_aigle_b200.synthetic import synthetic_raster
    wpath = str(tmp_path / "w.safetensors")
    bench.make_weights(wpath, seed=7)
    arr = synthetic_raster(700, 1000, seed=11)
    name = "mem://rescale_accumulate"
    register_raster(name, ZoneRaster(arr, L, T, RES, name=name))
    cfg = bench.zonal_config(wpath, str(tmp_path), name, 4)
    cfg["output_px_meters"] = 0.4
    cfg = inf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = cuda
    sizes = compute_patch_sizes(cfg)
    model = build_inference_model(cfg, sizes).to(cuda)
    tiles = generate_patches_from_reference(cfg, name, None)
    ds = inf.prep_dataset(cfg, tiles, sizes)
    canvas, transform = inf.inference(model, ds, tiles, cfg, name)
    assert tuple(canvas.shape) == (19, 350, 500) and tuple(transform)[0] == 0.4 and tuple(transform)[4] == -0.4
    cover = canvas.sum(dim=0)
    assert float(cover.min()) > 0.99                                   # every output pixel received a tile
    labels, conf = inf.logits_to_labels_and_confidence(canvas)
    RasterSink.write_files = False
    outs, _ = inf.init_outputs(cfg, name, 0)
    inf.inference_and_write(model, ds, tiles, cfg, outs, name)
    written = torch.from_numpy(outs[TASK].to_host()[0]).to(cuda)
    once = cover < 1.5
    assert float(once.float().mean()) > 0.5
    assert bool((labels[once] == written[once]).all())
    assert float(conf[once].max()) <= 1.0 + 1e-5

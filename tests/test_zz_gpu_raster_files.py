"""GPU: the zonal path FILE TO FILE -- GeoTIFF / JPEG-2000 inputs decoded by libfz_rasterio / OpenJPEG into page-locked memory
behind the upload, tiled LZW GeoTIFF / COG outputs, one zone file split over ranks.  These
tests live in a file that sorts last: they exercise the newest host code, and the kernel parity suites run before them."""
import os

import numpy as np
import pytest
import torch

from test_gpu_pipeline import L, RES, T, TASK, _no_tf32, _zone, setup  # noqa: F401  (fixtures and helpers of the pipeline tests)

pytestmark = pytest.mark.gpu


def test_geotiff_in_geotiff_out(setup, tmp_path):
    """The reference's file contract end to end: RGBI GeoTIFF on disk -> inference_and_write -> LZW GeoTIFF class raster
    with the input's georeferencing; identical to the run from the in-memory raster."""
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.geotiff import read_geotiff, write_geotiff
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    import bench
    tmp, wpath, _ = setup
    arr, cfg_mem = _zone(tmp, wpath, 1000, 700, 64, "mem://z_tif")
    # a 4-band GeoTIFF written the way GIS tools do (pixel interleaved): use Pillow directly
    from PIL import Image, TiffImagePlugin
    ifd = TiffImagePlugin.ImageFileDirectory_v2()
    ifd[33550] = (RES, RES, 0.0)
    ifd.tagtype[33550] = 12
    ifd[33922] = (0.0, 0.0, 0.0, L, T, 0.0)
    ifd.tagtype[33922] = 12
    src = str(tmp_path / "ortho.tif")
    Image.fromarray(np.ascontiguousarray(arr.transpose(1, 2, 0)), mode="RGBA").save(src, format="TIFF",
                                                                                    compression="tiff_lzw", tiffinfo=ifd)
    out_dir = str(tmp_path / "out")
    os.makedirs(out_dir)
    cfg = bench.zonal_config(wpath, out_dir, src, 4)
    cfg = inf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = torch.device("cuda:0")
    sizes = compute_patch_sizes(cfg)
    model = build_inference_model(cfg, sizes).to(cfg["device"])
    RasterSink.write_files = True
    results = {}
    for name, c, img in (("file", cfg, src), ("mem", cfg_mem, "mem://z_tif")):
        tiles = generate_patches_from_reference(c, img, None)
        ds = inf.prep_dataset(c, tiles, sizes)
        outs, _ = inf.init_outputs(c, img, 0)
        inf.inference_and_write(model, ds, tiles, c, outs, img)
        results[name] = (outs[TASK].to_host()[0].copy(), outs[TASK].written_path)
    assert np.array_equal(results["file"][0], results["mem"][0])
    got, left, top, res, _ = read_geotiff(results["file"][1])
    assert np.array_equal(got[0], results["file"][0]) and (left, top, res) == (L, T, RES)


def _run_inference_jp2_in_cog_out(setup, tmp_path):
    """The product script's file contract (scripts/run_fast_aigle_segmentation.py:75-119; inference.py:60 globs *.jp2):
    a JPEG-2000 ortho with a GeoJP2 box + a config file -> ``run_inference`` -> the class raster as a COG (``cog_conversion``,
    inference.py:633-641: LZW, 512 blocks, nearest overviews, the plain GeoTIFF removed), georeferenced like the input and
    identical to the run from the in-memory raster."""
    import json
    import bench
    from PIL import features
    if not features.check("jpg_2000"):
        pytest.skip("Pillow without OpenJPEG")
    from test_geotiff import _geojp2_box, _jp2_with_box
    from flair_for_aigle_b200 import raster_io
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    tmp, wpath, _ = setup
    arr, cfg_mem = _zone(tmp, wpath, 1000, 700, 64, "mem://z_jp2")
    src = str(tmp_path / "ortho.jp2")
    _jp2_with_box(src, np.ascontiguousarray(arr.transpose(1, 2, 0)), _geojp2_box(L, T, RES, 2154))     # lossless
    cfg = bench.zonal_config(wpath, str(tmp_path / "out"), src, 4)
    cfg["cog_conversion"] = True
    cfg_path = str(tmp_path / "zone.json")
    with open(cfg_path, "w") as f:
        json.dump(cfg, f)
    RasterSink.write_files = True
    written = inf.run_inference(cfg_path)
    cog = written[TASK]
    assert cog.endswith("_COG.tif") and os.path.isfile(cog) and not os.path.exists(cog.replace("_COG.tif", ".tif"))
    got, info = raster_io.read_raster(cog)
    assert info.tiled and info.block_w == 512 and info.compression == raster_io.COMP_LZW and info.overviews == 1
    assert (info.left, info.top, info.res_x, info.epsg) == (L, T, RES, 2154)
    sizes = compute_patch_sizes(cfg_mem)
    model = build_inference_model(cfg_mem, sizes).to(cfg_mem["device"])
    tiles = generate_patches_from_reference(cfg_mem, "mem://z_jp2", None)
    ds = inf.prep_dataset(cfg_mem, tiles, sizes)
    outs, _ = inf.init_outputs(cfg_mem, "mem://z_jp2", 0)
    inf.inference_and_write(model, ds, tiles, cfg_mem, outs, "mem://z_jp2")
    assert np.array_equal(got[0], outs[TASK].to_host()[0])
    assert np.array_equal(raster_io.read_raster(cog, level=1)[0][0], got[0][::2, ::2])        # nearest overview, even sizes



def test_run_inference_jp2_in_cog_out_pillow_decoder(setup, tmp_path, monkeypatch):
    """JPEG 2000 decoded by Pillow in one go (FZ_JP2_DECODER=pillow): the path that ran on the B200 first."""
    monkeypatch.setenv("FZ_JP2_DECODER", "pillow")
    _run_inference_jp2_in_cog_out(setup, tmp_path)


def test_zone_file_sharded_over_ranks_equals_single_run(setup, tmp_path):
    """SURVEY 8(e) through the public API: ``shard_zone`` + ``run_zone_shard`` (what ``run_inference`` does per rank under
    torch.distributed) on a GeoTIFF zone -- every rank decodes only its row strip of the file, runs it through
    ``inference_and_write`` and returns the class-raster rows it owns; the ranks' rows stacked are the single-run raster bit
    for bit.  (The ranks run one after the other on this GPU; tests/test_zone_shards.py moves the strips between gloo
    processes with the job's send / recv.)"""
    import bench
    from flair_for_aigle_b200 import raster_io
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    from flair_for_aigle_b200.synthetic import synthetic_raster
    tmp, wpath, _ = setup
    arr = synthetic_raster(1500, 1000, seed=13)
    src = str(tmp_path / "zone.tif")
    raster_io.write_geotiff(src, arr, L, T, RES, epsg=2154, pixel_interleave=True, predictor=2, block=256)
    out_dir = str(tmp_path / "out")
    os.makedirs(out_dir)
    cfg = bench.zonal_config(wpath, out_dir, src, 4)
    cfg = inf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = torch.device("cuda:0")
    sizes = compute_patch_sizes(cfg)
    model = build_inference_model(cfg, sizes).to(cfg["device"])
    tiles = generate_patches_from_reference(cfg, src, None)
    RasterSink.write_files = False
    try:
        ds = inf.prep_dataset(cfg, tiles, sizes)
        outs, _ = inf.init_outputs(cfg, src, 0)
        inf.inference_and_write(model, ds, tiles, cfg, outs, src)
        whole = outs[TASK].to_host()[0].copy()
        assert whole.shape == (1500, 1000) and whole.max() < 19
        import gc
        from flair_for_aigle_b200.flair_zonal_detection.raster import open_raster
        del ds, outs
        gc.collect()
        assert not open_raster(src).loaded                   # nobody holds the decoded zone any more: ranks start from the file
        for world in (2, 3):
            parts = []
            for rank in range(world):
                sh = inf.shard_zone(cfg, tiles, rank, world)
                owned = inf.run_zone_shard(model, sh, sizes)[TASK]
                assert tuple(owned.shape) == (1, sh.out_rows[1] - sh.out_rows[0], 1000)
                parts.append(owned[0].cpu().numpy())
            assert np.array_equal(np.concatenate(parts), whole), world
    finally:
        RasterSink.write_files = True


def test_zz_run_inference_jp2_in_cog_out_openjpeg_direct(setup, tmp_path, monkeypatch):
    """The same with the direct OpenJPEG binding (flair_for_aigle_b200/openjpeg.py): worker threads, decoded progressively
    behind the upload.  Written after the round's last GPU minutes, hence the LAST test of the GPU suite."""
    monkeypatch.delenv("FZ_JP2_DECODER", raising=False)
    from flair_for_aigle_b200.flair_zonal_detection import geotiff
    _run_inference_jp2_in_cog_out(setup, tmp_path)
    assert geotiff.row_source(str(tmp_path / "ortho.jp2")) is not None          # the direct decoder did take this file

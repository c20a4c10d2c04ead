"""CPU: one zone FILE over several ranks through the public API (flair_zonal_detection/inference.py: shard_zone,
gather_row_strips).  The forward pass needs a GPU, so a stand-in with the same contract as
``inference_and_write`` -- boundless tile windows read from the rank's raster, a prediction that depends on the pixels AND on
the position inside the tile, the margin-cropped write under last-writer ownership -- runs here on the strips the real call
would get; what is pinned is everything around the forward: which rows a rank decodes, its tiles, its georeferencing, the
rows it owns, and that the assembled raster equals the single-process one bit for bit.  The world-2 test runs the ranks as
gloo processes and moves the strips with the same send / recv the GPU job uses over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch

L, T, RES = 700000.0, 6600000.0, 0.2
P, MARGIN = 512, 64


def _config(path, out_dir):
    import bench
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    cfg = bench.zonal_config("unused", out_dir, path, 4)
    cfg["margin"] = MARGIN
    return inf.initialize_geometry_and_resolutions(cfg)


def _zone_file(path, W, H, seed=3):
    from flair_for_aigle_b200 import raster_io as rio
    arr = np.random.default_rng(seed).integers(0, 256, (4, H, W), dtype=np.uint8)
    rio.write_geotiff(path, arr, L, T, RES, epsg=2154, pixel_interleave=True, block=256)
    return arr


def _stand_in_inference_and_write(cfg, tiles, raster):
    """inference.py:254-355 with a toy 'model': class = (sum of the 4 bands + 3 * row in tile + 7 * column in tile) mod 19."""
    from flair_for_aigle_b200.flair_zonal_detection.slicing import ownership_windows, tile_plan
    plan = tile_plan(tiles, cfg["image_bounds"], cfg["reference_resolution"], P, MARGIN)
    own = ownership_windows(plan)
    arr = raster.read()
    C, H, W = arr.shape
    out = np.full((H, W), 255, np.uint8)
    yy, xx = np.mgrid[0:P, 0:P]
    for (r0, c0, top, left, h, w), (a, b, c, d) in zip(plan, own):
        if h == 0 or b <= a or d <= c:
            continue
        win = np.zeros((C, P, P), np.int64)
        rr0, rr1, cc0, cc1 = max(r0, 0), min(r0 + P, H), max(c0, 0), min(c0 + P, W)
        win[:, rr0 - r0:rr1 - r0, cc0 - c0:cc1 - c0] = arr[:, rr0:rr1, cc0:cc1]
        pred = ((win.sum(0) + 3 * yy + 7 * xx) % 19).astype(np.uint8)[MARGIN:P - MARGIN, MARGIN:P - MARGIN]
        out[a:b, c:d] = pred[a - top:b - top, c - left:d - left]
    return out


def test_shards_of_a_zone_file_reassemble_to_the_single_run(tmp_path, monkeypatch):
    from flair_for_aigle_b200 import raster_io as rio
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.raster import open_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    W, H = 1100, 3000
    path = str(tmp_path / "zone.tif")
    arr = _zone_file(path, W, H)
    cfg = _config(path, str(tmp_path))
    tiles = generate_patches_from_reference(cfg, path, None)
    whole = _stand_in_inference_and_write(cfg, tiles, open_raster(path))
    assert whole.max() < 19                                              # every pixel owned by somebody
    windows = []
    real = rio.read_window
    monkeypatch.setattr(rio, "read_window", lambda p, r0, c0, h, w, **k: (windows.append((r0, h)), real(p, r0, c0, h, w, **k))[1])
    for world in (1, 2, 3, 8, 12):
        assembled = np.full((H, W), 255, np.uint8)
        n_tiles, covered = 0, np.zeros(H, np.int32)
        for rank in range(world):
            del windows[:]
            sh = inf.shard_zone(cfg, tiles, rank, world)
            assert sh.all_out_rows == inf.shard_zone(cfg, tiles, 0, world).all_out_rows and sh.all_out_rows[rank] == sh.out_rows
            if sh.raster is None:                                        # more ranks than tile rows (8 tile rows here)
                assert world > 8 and len(sh.tiles) == 0 and sh.out_rows == (0, 0) and inf.run_zone_shard(None, sh, {}) == {}
                continue
            (i0, i1), (o0, o1) = sh.in_rows, sh.out_rows
            assert i0 <= o0 < o1 <= i1 and (sh.raster.height, sh.raster.width) == (i1 - i0, W) and not sh.raster.loaded
            assert sh.raster.top == T - i0 * RES and sh.raster.left == L and sh.raster.crs == "EPSG:2154"
            assert sh.config["image_bounds"]["top"] == sh.raster.top and sh.config["image_shape_px"] == {"height": i1 - i0, "width": W}
            assert windows == []                                          # sharding itself decodes nothing
            assert np.array_equal(sh.raster.read(), arr[:, i0:i1])
            assert windows == [(i0, i1 - i0)]                             # ... and the rank decodes its rows only, once
            strip_out = _stand_in_inference_and_write(sh.config, sh.tiles, sh.raster)
            assembled[o0:o1] = strip_out[o0 - i0:o1 - i0]
            covered[o0:o1] += 1
            n_tiles += len(sh.tiles)
        assert n_tiles == len(tiles) and (covered == 1).all()
        assert np.array_equal(assembled, whole), world
    # config untouched; an in-memory raster shards as views of the same pixels
    assert cfg["modalities"]["AERIAL_RGBI"]["input_img_path"] == path
    whole_r = open_raster(path)
    whole_r.read()
    strip = whole_r.row_strip(100, 900)
    assert np.shares_memory(strip.read(), whole_r.read()) and strip.top == T - 100 * RES and strip.height == 800
    with pytest.raises(ValueError, match="empty row strip"):
        whole_r.row_strip(50, 50)
    cfg2 = dict(cfg)
    cfg2["output_px_meters"] = 0.4
    with pytest.raises(NotImplementedError, match="output_px_meters"):
        inf.shard_zone(cfg2, tiles, 0, 2)


def test_progressive_strip_decode(tmp_path):
    """A rank's strip decodes progressively too: bottom rows first, slabs on the FILE's block rows."""
    from flair_for_aigle_b200.flair_zonal_detection.raster import open_raster
    path = str(tmp_path / "zone.tif")
    arr = _zone_file(path, 600, 3000)
    strip = open_raster(path).row_strip(700, 2900)
    prog = strip.begin_progressive()
    assert prog is not None and (prog.row0, prog.rows) == (700, 2200)
    prog.wait_rows(2000)
    assert np.array_equal(prog.array[:, 2000:], arr[:, 2700:2900])
    prog.wait_all()
    assert prog.lo == 0 and np.array_equal(strip.read(), arr[:, 700:2900])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, out_dir, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.raster import open_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    cfg = _config(path, out_dir)
    tiles = generate_patches_from_reference(cfg, path, None)
    sh = inf.shard_zone(cfg, tiles, rank, world)
    (i0, _), (o0, o1) = sh.in_rows, sh.out_rows
    strip_out = _stand_in_inference_and_write(sh.config, sh.tiles, sh.raster)
    owned = torch.from_numpy(strip_out[None, o0 - i0:o1 - i0].copy())
    ref = open_raster(path)
    full = torch.full((1, ref.height, ref.width), 255, dtype=torch.uint8) if rank == 0 else None
    inf.gather_row_strips(owned, sh.all_out_rows, full, rank, world)
    if rank == 0:
        whole = _stand_in_inference_and_write(cfg, tiles, ref)
        q.put(bool(np.array_equal(full[0].numpy(), whole)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_row_strips_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    path = str(tmp_path / "zone.tif")
    _zone_file(path, 900, 2600, seed=5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(60)
    assert ok and all(p.exitcode == 0 for p in procs)

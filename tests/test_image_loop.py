"""CPU: the product script's image loop (flair_for_aigle_b200/scripts/run_fast_aigle_segmentation.py <- reference
scripts/run_fast_aigle_segmentation.py:82-132).  The three GPU calls -- init_outputs, inference_and_write,
raster_to_polygons -- are replaced by stand-ins; what is checked is the loop around them: per-image config re-pointing,
geozone misses, results already on disk skipped (the script's resume rule), one GeoPackage per image, the aggregation, and
the host pipelining (the next image is already decoding while the current one is 'on the GPU')."""
import os

import numpy as np

from flair_for_aigle_b200 import raster_io as rio

L, T, RES = 700000.0, 6600000.0, 0.2


def _square(x0, y0, s):
    return [[x0, y0], [x0 + s, y0], [x0 + s, y0 + s], [x0, y0 + s], [x0, y0]]


def test_image_loop_resume_prefetch_and_aggregation(tmp_path, monkeypatch):
    import bench
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.polygonize import PolygonTable
    from flair_for_aigle_b200.flair_zonal_detection.raster import open_raster
    from flair_for_aigle_b200.scripts import run_fast_aigle_segmentation as script
    rng = np.random.default_rng(0)
    folder, results = tmp_path / "images", tmp_path / "results"
    folder.mkdir()
    images = []
    for k in range(4):                                        # four orthos side by side, 200 m apart
        p = str(folder / f"ortho_{k}.tif")
        rio.write_geotiff(p, rng.integers(0, 256, (4, 700, 1000), dtype=np.uint8), L + 200.0 * k, T, RES, epsg=2154,
                          pixel_interleave=True, block=256)
        images.append(p)
    weights = tmp_path / "w.safetensors"
    weights.write_bytes(b"x")
    cfg = bench.zonal_config(str(weights), str(results), images[0], 4)
    cfg = inf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = "cpu"
    seen = []

    class Sink:
        def __init__(self, name):
            self.name = name

    def fake_init_outputs(config, ref_img, i=0):
        return {"task": Sink(open_raster(ref_img).name)}, {"task": "unused"}

    def fake_inference_and_write(model, dataset, tiles, config, outs, ref_img):
        path = open_raster(ref_img).name
        k = images.index(path)
        nxt = images[k + 1] if k + 1 < len(images) else None
        # the NEXT image's decode has been started by the loop (somebody holds it, so open_raster returns that raster)
        started = nxt is not None and open_raster(nxt)._progress is not None
        b = config["image_bounds"]
        seen.append((k, len(tiles), started, b["left"], config["modalities"]["AERIAL_RGBI"]["input_img_path"] == path))
        assert dataset.readers["AERIAL_RGBI"] is open_raster(path)

    def fake_polygons(output_files, n_jobs=None):
        k = images.index(output_files["task"].name)
        if k == 2:                                            # an image without any kept polygon: no file (:121)
            return PolygonTable(np.zeros(0, np.int64), np.zeros(0), [], "EPSG:2154")
        geoms = [{"type": "Polygon", "coordinates": [_square(L + 200.0 * k + 10 * j, T - 50, 5)]} for j in range(k + 1)]
        return PolygonTable(np.full(k + 1, 6 + k), np.full(k + 1, 25.0), geoms, "EPSG:2154")
    monkeypatch.setattr(inf, "init_outputs", fake_init_outputs)
    monkeypatch.setattr(inf, "inference_and_write", fake_inference_and_write)
    monkeypatch.setattr(inf, "raster_to_polygons", fake_polygons)

    class Zone:                                               # a contour covering images 0..2 only
        bounds = (L, T - 140.0, L + 590.0, T)
    written = script.segment_images(None, cfg, images, str(results), Zone(), patch_sizes={"AERIAL_RGBI": 512})
    assert [os.path.basename(w) for w in written] == ["ortho_0.gpkg", "ortho_1.gpkg"]
    assert [s[0] for s in seen] == [0, 1, 2]                  # image 3 lies outside the zone: sliced into 0 tiles, skipped
    assert all(s[1] == 6 for s in seen[:2]) and seen[2][1] > 0
    assert [s[2] for s in seen] == [True, True, True]         # prefetch: the following image was already decoding
    assert [s[3] for s in seen] == [L, L + 200.0, L + 400.0] and all(s[4] for s in seen)
    assert sorted(os.listdir(results)) == ["ortho_0.gpkg", "ortho_1.gpkg"]
    # resume (:92-95): results on disk are not recomputed
    del seen[:]
    again = script.segment_images(None, cfg, images, str(results), Zone(), patch_sizes={"AERIAL_RGBI": 512}, prefetch=False)
    assert again == [] and [s[0] for s in seen] == [2] and seen[0][2] is False
    table = script.aggregate_results(str(results))            # :131-132
    assert len(table) == 3 and table.class_id.tolist() == [6, 7, 7] and table.crs == "EPSG:2154"
    assert table.geometry[2]["coordinates"][0][0] == [L + 210.0, T - 50]
    assert script.result_path("/r", "/data/x/ortho_5.jp2") == "/r/ortho_5.gpkg"

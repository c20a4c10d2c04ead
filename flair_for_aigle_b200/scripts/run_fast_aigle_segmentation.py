"""The image loop of the product script: drop-in for the part of scripts/run_fast_aigle_segmentation.py that drives the
zonal path (:82-132) -- every ortho of a folder through slicing -> dataset -> ``inference_and_write`` ->
``raster_to_polygons`` -> one GeoPackage per image, images with a result on disk skipped (the script's resume rule,
:92-95), then all per-image files read back and concatenated.

What surrounds that loop in the reference -- S3 download, the geozone database, the SQL export, class remapping for the
Aigle application (:25-80, :135-end; utils/) -- is control plane and stays out of scope: the caller hands in
the config, the model checkpoint, the image folder and, optionally, the zone contour's bounding geometry.

Host pipelining the reference does not have: with JPEG-2000 orthos the decode takes seconds against a fraction of a second of
GPU work per image, so while image i is on the GPU the decode of image i + 1 already runs (``ZoneRaster.begin_progressive``:
OpenJPEG / libfz_rasterio worker threads), and the GeoPackage of image i - 1 is written by a background thread.
"""
from __future__ import annotations

import logging
import os
import threading
import time
from typing import List, Optional

from ..flair_zonal_detection import inference as inf
from ..flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
from ..flair_zonal_detection.polygonize import PolygonTable
from ..flair_zonal_detection.raster import open_raster
from ..flair_zonal_detection.slicing import generate_patches_from_reference

logger = logging.getLogger(__name__)


def result_path(result_folder: str, source_image_path: str) -> str:
    """scripts/run_fast_aigle_segmentation.py:90: ``<result_folder>/<image name>.gpkg``."""
    return os.path.join(result_folder, os.path.splitext(os.path.basename(source_image_path))[0] + ".gpkg")


def _prefetch(path: Optional[str]):
    """Open the next image and start decoding it in the background; the returned raster must be kept alive by the caller
    (open_raster shares a file's decode between everybody who holds it)."""
    if path is None:
        return None
    try:
        raster = open_raster(path)
        raster.begin_progressive()
        return raster
    except Exception as e:  # noqa: BLE001 -- the image's own turn reports the problem
        logger.warning(f"prefetch of {path} failed: {e!r}")
        return None


def segment_images(model, model_config_args: dict, images: List[str], result_folder: str, geozone_geometry_contour=None,
                   patch_sizes: Optional[dict] = None, n_jobs: int = 4, prefetch: bool = True) -> List[str]:
    """scripts/run_fast_aigle_segmentation.py:86-125.  -> the per-image GeoPackage paths written by THIS call (images whose
    result exists already, or whose polygons are empty, contribute none -- like the reference's ``global_results``)."""
    patch_sizes = patch_sizes or compute_patch_sizes(model_config_args)
    ref_mod = model_config_args['reference_modality']
    os.makedirs(result_folder, exist_ok=True)
    todo = []
    for source_image_path in images:
        out = result_path(result_folder, source_image_path)
        if os.path.exists(out):                                                          # :92-95 check if already run
            logger.warning(f"intermediate result found : {out} - raster skipped : {os.path.basename(source_image_path)}")
            continue
        todo.append((source_image_path, out))
    global_results: List[str] = []
    writer: Optional[threading.Thread] = None
    writer_error: List[BaseException] = []
    held = _prefetch(todo[0][0]) if (prefetch and todo) else None
    for i, (source_image_path, out) in enumerate(todo):
        start = time.time()
        current = held                                                                   # keeps this image's decode alive
        model_config_args['modalities'][ref_mod]['input_img_path'] = source_image_path   # :100
        model_config_args.pop('image_shape_px', None)
        model_config_args = inf.initialize_geometry_and_resolutions(model_config_args)   # :101
        tiles_gdf = generate_patches_from_reference(model_config_args, source_image_path, geozone_geometry_contour)
        logger.info(f"[✓] {source_image_path} Sliced into {len(tiles_gdf)} tiles in {time.time() - start:.2f}s")
        held = _prefetch(todo[i + 1][0]) if (prefetch and i + 1 < len(todo)) else None   # decodes while this image computes
        if len(tiles_gdf) == 0:                                                          # :106 the zone misses this image
            del current
            continue
        dataset = inf.prep_dataset(model_config_args, tiles_gdf, patch_sizes)
        ref_img = open_raster(source_image_path)
        output_files, _ = inf.init_outputs(model_config_args, ref_img, i)
        inf.inference_and_write(model, dataset, tiles_gdf, model_config_args, output_files, ref_img)
        gdf_results = inf.raster_to_polygons(output_files, n_jobs=n_jobs)                # :119
        del dataset, ref_img, current
        if len(gdf_results) > 0:                                                         # :121-124
            if writer is not None:
                writer.join()

            def write(table=gdf_results, path=out):
                try:
                    table.to_file(path, driver="GPKG")
                except BaseException as e:  # noqa: BLE001
                    writer_error.append(e)
            writer = threading.Thread(target=write, name="fz-gpkg-writer")
            writer.start()
            global_results.append(out)
        logger.info(f"[✓] Inference completed in {time.time() - start:.2f}s")
    if writer is not None:
        writer.join()
    if writer_error:
        raise writer_error[0]
    return global_results


def aggregate_results(result_folder: str) -> PolygonTable:
    """:131-132: every ``*.gpkg`` of the result folder read back and concatenated (sorted by name: ``os.listdir`` order is
    arbitrary)."""
    files = sorted(f for f in os.listdir(result_folder) if f.endswith('.gpkg'))
    return PolygonTable.concat([PolygonTable.read_file(os.path.join(result_folder, f)) for f in files])


def run_fast_aigle_segmentation(model_config_path: str, model_ckpt_path: str, images_folder: str, result_folder: str,
                                log_folder: Optional[str] = None, model_threshold_filepath: Optional[str] = None,
                                geozone_geometry_contour=None) -> PolygonTable:
    """The reference's entry point takes a namespace of S3 / database settings; here the same work from its local inputs:
    config (:74), model (:81), the folder's images (:86; ``*.jp2`` first, then GeoTIFFs), the per-image loop, the aggregation."""
    start_total = time.time()
    model_config_args = inf.prep_config(model_config_path, model_ckpt_path, model_threshold_filepath, result_folder,
                                        log_folder or result_folder, images_folder)
    patch_sizes = compute_patch_sizes(model_config_args)
    model = build_inference_model(model_config_args, patch_sizes).to(model_config_args['device'])
    images = inf.list_rasters(images_folder)
    segment_images(model, model_config_args, images, result_folder, geozone_geometry_contour, patch_sizes)
    logger.info(f"\n[✓] Total time: {time.time() - start_total:.2f}s")
    return aggregate_results(result_folder)

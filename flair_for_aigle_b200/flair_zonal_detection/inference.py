"""Zonal inference pipeline: drop-in for flair_zonal_detection/inference.py.

Same entry points and call order as the reference's working caller
(scripts/run_fast_aigle_segmentation.py:75-119):

    config  = prep_config(...) / initialize_geometry_and_resolutions(config)
    sizes   = compute_patch_sizes(config)
    model   = build_inference_model(config, sizes).to(config['device'])
    tiles   = generate_patches_from_reference(config, img_path, geozone)
    dataset = prep_dataset(config, tiles, sizes);  loader = DataLoader(dataset, batch_size=...)
    outs, _ = init_outputs(config, ref_img, i)
    inference_and_write(model, loader, tiles, config, outs, ref_img)

What changes is where the work happens: the raster stays in HBM as uint8, tiles are gathered by
a kernel, the encoder/decoder run on tcgen05, and the margin crop + argmax + windowed write of
inference.py:297-352 is the head convolution's epilogue.  No logits cross PCIe.
"""
from __future__ import annotations

import glob
import logging
import os
import time
from typing import Dict, NamedTuple, Optional, Tuple

import numpy as np
import torch

from .. import native as nv
from ..engine.zonal import ZonalRunner
from .config import config_recap_1, config_recap_2, load_config, validate_config
from .dataset import MultiModalSlicedDataset, normalization_affine
from .model_utils import build_inference_model, compute_patch_sizes
from .postprocess import convert, convert_to_cog  # noqa: F401  (re-exported like the reference)
from .raster import RasterSink, ZoneRaster, open_raster
from .slicing import generate_patches_from_reference, ownership_windows, tile_plan

logger = logging.getLogger(__name__)

RASTER_GLOBS = ("*.jp2", "*.tif", "*.tiff", "*.npy")


def overwrite_config(config, model_ckpt_path, model_threshold_filepath, result_folder, log_folder) -> Dict:
    """inference.py:42-51."""
    config['model_weights'] = model_ckpt_path
    config['model_threshold_filepath'] = model_threshold_filepath
    config['output_path'] = result_folder
    config['log_folder'] = log_folder
    return config


def list_rasters(images_folder: str):
    files = []
    for pat in RASTER_GLOBS:
        files += sorted(glob.glob(os.path.join(images_folder, pat)))
    return files


def prep_config(config_path: str, model_ckpt_path: str = None, model_threshold_filepath: str = None,
                result_folder: str = None, log_folder: str = None, images_folder: str = None) -> Dict:
    """inference.py:54-73.  The five override arguments are optional here so that the intent of
    the reference's stale ``run_inference(config_path)`` also works."""
    config = load_config(config_path)
    if images_folder is not None:
        rasters = list_rasters(images_folder)
        if not rasters:
            raise FileNotFoundError(f"no raster found in {images_folder}")
        config['modalities']['AERIAL_RGBI']['input_img_path'] = rasters[0]
    if model_ckpt_path is not None:
        config = overwrite_config(config, model_ckpt_path, model_threshold_filepath, result_folder, log_folder)
    validate_config(config)
    config.setdefault('output_type', 'argmax')
    config_recap_1(config)
    config = initialize_geometry_and_resolutions(config)
    config_recap_2(config)
    config['device'] = torch.device("cuda" if config.get("use_gpu", torch.cuda.is_available()) else "cpu")
    if config['device'].type == "cuda" and os.environ.get("LOCAL_RANK", "").isdigit():
        config['device'] = torch.device("cuda", int(os.environ["LOCAL_RANK"]))       # one process per GPU (torchrun)
    config['output_type'] = config.get("output_type", "argmax")
    return config


def initialize_geometry_and_resolutions(config: Dict) -> Dict:
    """inference.py:76-132 (same keys set, same bounds check, same "smallest m/px" reference rule)."""
    modalities = config['modalities']
    active = [m for m, on in modalities['inputs'].items() if on]
    resolutions, bounds = {}, []
    for mod in active:
        src = open_raster(modalities[mod]['input_img_path'])
        resolutions[mod] = round(src.res[0], 5)
        bounds.append((mod, src.bounds))
        if 'image_shape_px' not in config:
            config['image_shape_px'] = {'height': src.height, 'width': src.width}
    ref_mod, ref_bounds = bounds[0]
    for mod, b in bounds[1:]:
        if not np.allclose(b, ref_bounds, atol=1e-2):
            raise ValueError(f"[✗] Bounds mismatch between '{ref_mod}' and '{mod}':\n  {ref_mod}: {ref_bounds}\n"
                             f"  {mod}: {b}")
    ref_mod, reference_resolution = min(resolutions.items(), key=lambda kv: kv[1])
    config['reference_modality'] = ref_mod
    config['reference_resolution'] = reference_resolution
    config['modality_resolutions'] = resolutions
    config['image_bounds'] = {'left': ref_bounds.left, 'bottom': ref_bounds.bottom, 'right': ref_bounds.right,
                              'top': ref_bounds.top}
    config['tile_size_m'] = round(config['img_pixels_detection'] * reference_resolution, 2)
    config['margin_size_m'] = round(config['margin'] * reference_resolution, 2)
    return config


def prep_dataset(config: Dict, tiles_gdf, patch_sizes: Dict[str, int]) -> MultiModalSlicedDataset:
    """inference.py:136-154."""
    active = [m for m, on in config['modalities']['inputs'].items() if on]
    modality_cfgs = {m: config['modalities'][m] for m in active}
    config['labels'] = [t['name'] for t in config['tasks'] if t['active']]
    config['labels_configs'] = {t['name']: {'value_name': t['class_names']} for t in config['tasks'] if t['active']}
    return MultiModalSlicedDataset(dataframe=tiles_gdf, modality_cfgs=modality_cfgs, patch_size_dict=patch_sizes,
                                   ref_date_str=config.get('multitemp_model_ref_date'), modalities_config=config)


def init_outputs(config: Dict, ref_img, i=0) -> Tuple[Dict[str, RasterSink], Dict[str, str]]:
    """inference.py:157-208: one uint8 output raster per active task (1 band for argmax, n_cls
    bands for class_prob), rescaled grid when ``output_px_meters`` differs."""
    ref_img = open_raster(ref_img)
    output_files, temp_paths = {}, {}
    output_type = config['output_type']
    ref_res = config['reference_resolution']
    out_res = config.get("output_px_meters", ref_res)
    ib = config['image_bounds']
    needs_rescale = abs(ref_res - out_res) > 1e-6
    device = config.get('device', torch.device('cuda'))
    for task in config['tasks']:
        if not task['active']:
            continue
        n_cls = len(task['class_names'])
        suffix = 'argmax' if output_type == 'argmax' else 'class-prob'
        out_path = os.path.join(config['output_path'], f"{config['output_name']}_{task['name']}_{suffix}_i.tif")
        if not needs_rescale:
            h, w = ref_img.height, ref_img.width
        else:
            h = int(round((ib['top'] - ib['bottom']) / out_res))
            w = int(round((ib['right'] - ib['left']) / out_res))
        output_files[task['name']] = RasterSink(out_path, n_cls if output_type == "class_prob" else 1, h, w,
                                                ib['left'], ib['top'], out_res, ref_img.crs, device=device)
        temp_paths[task['name']] = out_path
    return output_files, temp_paths


def resample_prediction(prediction, scale: float):
    """inference.py:212-226: nearest-neighbour zoom of a (H, W) or (C, H, W) prediction (scipy.ndimage.zoom, order 0),
    as separable index gathers.  Host helper with the reference's signature; the zonal path applies the same index
    map inside the crop kernels (fz_crop_zoom_write)."""
    from .slicing import zoom_map
    if abs(scale - 1.0) < 1e-9:
        return prediction
    prediction = np.asarray(prediction)
    if prediction.ndim not in (2, 3):
        raise ValueError(f"Unexpected prediction shape: {prediction.shape}")
    zy, zx = zoom_map(prediction.shape[-2], scale), zoom_map(prediction.shape[-1], scale)
    out = prediction[..., np.maximum(zy, 0), :][..., np.maximum(zx, 0)]
    out[..., zy < 0, :] = 0          # scipy's constant fill (see zoom_map)
    out[..., zx < 0] = 0
    return out


def _rescale(config):
    """(needs_rescale, scale) of inference.py:299-303."""
    ref_res = config['reference_resolution']
    out_res = config.get('output_px_meters', ref_res)
    needs = abs(ref_res - out_res) > 1e-6
    return needs, (ref_res / out_res if needs else 1.0)


def _runner(model, config, margin: int) -> ZonalRunner:
    key = "_zonal_runner"
    task = config['labels'][0] if 'labels' in config else [t['name'] for t in config['tasks'] if t['active']][0]
    eng = model.engine(task, max_batch=int(config.get('batch_size', model.max_batch)))
    r = getattr(model, key, None)
    if r is None or r.eng is not eng or r.margin != margin:
        r = ZonalRunner(eng, margin, use_graph=bool(config.get('use_cuda_graph', True)), norm=model._norm)
        setattr(model, key, r)
    return r


@torch.no_grad()
def inference_and_write(model, dataloader, tiles_gdf, config: Dict, output_files: Dict[str, RasterSink],
                        ref_img) -> None:
    """inference.py:254-355.  Writes every tile's margin-cropped prediction into the task's output
    raster (later tiles overwrite earlier ones) and closes the rasters."""
    device = torch.device(config['device'])
    if device.type != "cuda":
        raise nv.NativeError("inference_and_write runs on CUDA only (no CPU fallback)")
    margin = int(config['margin'])
    P = int(config['img_pixels_detection'])
    output_type = config['output_type']
    needs_rescale, scale = _rescale(config)
    ref_img = open_raster(ref_img)
    b = ref_img.bounds
    ib = {'left': b.left, 'bottom': b.bottom, 'right': b.right, 'top': b.top}
    ref_res = config['reference_resolution']
    plan = tile_plan(tiles_gdf, ib, ref_res, P, margin, config.get('output_px_meters', ref_res))
    own = ownership_windows(plan)
    dataset = getattr(dataloader, 'dataset', dataloader)
    tasks = [t['name'] for t in config['tasks'] if t['active']]
    zmap_d = None
    if needs_rescale:
        from .slicing import zoom_map
        zmap_d = torch.from_numpy(zoom_map(P - 2 * margin, scale)).to(device)

    if (isinstance(dataset, MultiModalSlicedDataset) and output_type == "argmax" and len(tasks) == 1
            and not needs_rescale and len(model.active_mono) == 1):
        # fused device path: feeder -> encoder/decoder -> head epilogue writes the class raster
        mod = model.active_mono[0]
        sink = output_files[tasks[0]]
        runner = _runner(model, config, margin)
        host = dataset.host_raster(mod)
        if host.is_pinned() and bool(config.get('stream_upload', True)):
            # upload and read-back overlapped with the forward
            runner.run_streamed(host, plan, own, sink.device_array[0], out_host=sink.pinned_buffer()[0],
                                rows_ready=dataset.host_rows_ready(mod))
            sink.mark_streamed()
        else:
            runner.run(dataset.device_raster(mod, device), plan, own, sink.device_array[0])
    else:
        # generic path (any iterable of reference-style batches, class_prob output, several tasks):
        # model(inputs) -> logits stay on the device -> crop/convert/write kernels
        plan_d = torch.from_numpy(plan).to(device)
        own_d = torch.from_numpy(own).to(device)
        for batch in _iter_batches(dataloader, dataset, model, config, device):
            idx = batch.pop('index').to(device).flatten().long()
            logits_tasks, _ = model(batch)
            for task, logits in logits_tasks.items():
                sink = output_files[task]
                pl, ow = plan_d[idx].contiguous(), own_d[idx].contiguous()
                if needs_rescale:
                    # inference.py:303-312: argmax first, then zoom the labels; class_prob zooms the logits first --
                    # with a nearest-neighbour zoom both are the same gather, done inside the crop kernel
                    nv.crop_zoom_write(0 if output_type == "argmax" else 1, logits, nv.NCHW, margin, pl, ow, zmap_d,
                                       sink.device_array[0] if output_type == "argmax" else sink.device_array)
                elif output_type == "argmax":
                    nv.crop_argmax_write(logits, nv.NCHW, margin, pl, ow, sink.device_array[0])
                else:
                    nv.crop_softmax_write(logits, nv.NCHW, margin, pl, ow, sink.device_array)
    torch.cuda.synchronize(device)
    for dst in output_files.values():
        dst.close()


from .polygonize import (PolygonTable, raster_to_polygons, vectorize_segmentation,  # noqa: E402,F401  (inference.py:375-407,
                         vectorize_segmentation_parallel)                            # :574-632)

def _iter_batches(dataloader, dataset, model, config, device):
    """Reference-style batches ({MOD: (B,C,P,P) fp32 normalised, 'index': ...}).  For our own
    dataset the tensors are produced on the device by the feeder kernel."""
    if not isinstance(dataset, MultiModalSlicedDataset):
        for batch in dataloader:
            yield {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items()
                   if not k.endswith('_RAW')}
        return
    # every active mono-temporal modality: its own raster, window origins in its own pixel grid and normalisation
    # (dataset.py:174-209); a single modality is FusionHandler case 1, two or more go through conv_f (flair_model.py:473-547)
    feeds = []
    for mod in model.active_mono:
        raster = dataset.device_raster(mod, device)
        C = raster.shape[0]
        means, stds = normalization_affine(dataset.modalities[mod].get('normalization'), C,
                                           np.uint8 if raster.dtype == torch.uint8 else np.float32)
        mean = torch.tensor(means, dtype=torch.float32, device=device)
        std = torch.tensor(stds, dtype=torch.float32, device=device)
        kind, rplan = dataset.modality_read_plan(mod)      # whole-pixel copy, or rasterio's resampled (bilinear) read
        gather = nv.gather_tiles_f32 if kind == "aligned" else nv.gather_tiles_resampled
        feeds.append((mod, raster, mean, std, torch.from_numpy(rplan).to(device), gather,
                      int(dataset.patch_sizes.get(mod, config['img_pixels_detection']))))
    bs = int(config.get('batch_size', 8))
    for s in range(0, len(dataset), bs):
        idx = torch.arange(s, min(s + bs, len(dataset)), device=device)
        batch = {mod: gather(raster, rplan[idx].contiguous(), ps, mean, std)
                 for mod, raster, mean, std, rplan, gather, ps in feeds}
        batch['index'] = idx
        yield batch


@torch.no_grad()
def inference(model, dataloader, tiles_gdf, config: Dict, raster_img):
    """inference.py:468-564 with its intended semantics (wide accumulator instead of
    the wrapping int8 one, the window of inference.py:318-321): softmax of every margin-cropped
    tile accumulated into a (n_cls,H,W) float32 canvas on the device.  Returns (canvas, transform)."""
    device = torch.device(config['device'])
    if device.type != "cuda":
        raise nv.NativeError("inference runs on CUDA only (no CPU fallback)")
    margin = int(config['margin'])
    P = int(config['img_pixels_detection'])
    needs_rescale, scale = _rescale(config)
    raster_img = open_raster(raster_img)
    b = raster_img.bounds
    ib = {'left': b.left, 'bottom': b.bottom, 'right': b.right, 'top': b.top}
    ref_res = config['reference_resolution']
    out_res = config.get('output_px_meters', ref_res)
    plan_np = tile_plan(tiles_gdf, ib, ref_res, P, margin, out_res)
    plan = torch.from_numpy(plan_np).to(device)
    if needs_rescale:
        # inference.py:515-523,538-539: the zoomed tiles land on the out_res grid, whose size the reference computes at
        # :538-539 (its canvas keeps the input shape, which only fits when out_res >= ref_res); the canvas here IS that grid
        from .slicing import zoom_map
        H = int(round((ib['top'] - ib['bottom']) / out_res))
        W = int(round((ib['right'] - ib['left']) / out_res))
        zmap_d = torch.from_numpy(zoom_map(P - 2 * margin, scale)).to(device)
    else:
        H, W, zmap_d = raster_img.height, raster_img.width, None
    canvas = torch.zeros((model.task_nclasses, H, W), dtype=torch.float32, device=device)
    dataset = getattr(dataloader, 'dataset', dataloader)
    for batch in _iter_batches(dataloader, dataset, model, config, device):
        idx = batch.pop('index').to(device).flatten().long()
        logits_tasks, _ = model(batch)
        host_rows = plan_np[idx.cpu().numpy()]          # the batch's windows on the host: disjoint windows share a launch
        for _, logits in logits_tasks.items():
            if needs_rescale:
                nv.crop_zoom_accumulate(logits, nv.NCHW, margin, plan[idx].contiguous(), zmap_d, canvas, plan_host=host_rows)
            else:
                nv.crop_softmax_accumulate(logits, nv.NCHW, margin, plan[idx].contiguous(), None, canvas,
                                           plan_host=host_rows)
    torch.cuda.synchronize(device)
    transform = raster_img.profile['transform']
    if needs_rescale:   # same origin, out_res pixels (init_outputs' profile, inference.py:186-194)
        t = tuple(transform)
        transform = (out_res, t[1], t[2], t[3], -out_res, t[5])
    return canvas, transform


def logits_to_labels_and_confidence(probs):
    """inference.py:566-572 on the device canvas."""
    if not torch.is_tensor(probs):
        probs = torch.from_numpy(np.ascontiguousarray(probs))
    if not probs.is_cuda:
        if not torch.cuda.is_available():
            raise nv.NativeError("logits_to_labels_and_confidence runs on CUDA only (no CPU fallback)")
        probs = probs.cuda()
    return nv.canvas_argmax(probs.float().contiguous(), want_confidence=True)


# ---------------------------------------------------------------------------------------------------- one zone, several GPUs
class ZoneShard(NamedTuple):
    """One rank's part of a zone split into row strips (engine/strips.py)."""
    rank: int
    world: int
    config: Dict            # the zonal config re-pointed at this rank's strip raster(s)
    tiles: object           # this rank's rows of the GLOBAL tile table (zone coordinates)
    raster: ZoneRaster      # the reference modality's strip: input rows [in_rows) of the zone, nothing else is read
    in_rows: Tuple[int, int]    # zone rows the strip holds (output rows owned + the margin halo)
    out_rows: Tuple[int, int]   # zone rows of the class raster this rank OWNS
    all_out_rows: Tuple[Tuple[int, int], ...]   # the same for every rank (each rank derives them from the global plan)


def shard_zone(config: Dict, tiles_gdf, rank: int, world: int) -> ZoneShard:
    """The reference runs a zone on one device (inference.py:71).  Here the zone's tile ROWS are dealt to ``world`` ranks;
    rank r gets a config whose modality rasters are row strips of the input files -- only those rows are decoded
    (``ZoneRaster.row_strip`` -> ``fzio_read_window``) --, its rows of the global tile table, and the class-raster rows it
    owns under the last-writer rule.  There is no data-path collective: neighbouring ranks read the same halo rows from the
    file, and the union of the owned rows is the single-GPU raster bit for bit.  When ``world`` exceeds the number of tile
    rows the surplus ranks get an empty shard (no tiles, no raster)."""
    from ..engine.strips import shard_rows
    needs_rescale, _ = _rescale(config)
    if needs_rescale:
        raise NotImplementedError("sharding a zone whose output_px_meters differs from the reference resolution")
    ref_mod = config['reference_modality']
    ref = open_raster(config['modalities'][ref_mod]['input_img_path'])
    P, margin, ref_res = int(config['img_pixels_detection']), int(config['margin']), config['reference_resolution']
    plan = tile_plan(tiles_gdf, config['image_bounds'], ref_res, P, margin, config.get('output_px_meters', ref_res))
    own = ownership_windows(plan)
    shards = shard_rows(plan, own, P, ref.height, world)
    me = shards[rank]
    all_rows = tuple((sh.out_r0, sh.out_r1) for sh in shards)
    if len(me.tile_idx) == 0:                               # more ranks than tile rows: nothing to do here
        return ZoneShard(rank, world, config, tiles_gdf.iloc[:0], None, (0, 0), (0, 0), all_rows)
    cfg = dict(config)
    cfg['modalities'] = {k: (dict(v) if isinstance(v, dict) else v) for k, v in config['modalities'].items()}
    strip = None
    for mod, on in config['modalities']['inputs'].items():
        if not on:
            continue
        src = open_raster(config['modalities'][mod]['input_img_path'])
        if (src.height, src.width) != (ref.height, ref.width):
            raise NotImplementedError(f"sharding with modality '{mod}' on another pixel grid than '{ref_mod}'")
        s = src.row_strip(me.in_r0, me.in_r1)
        cfg['modalities'][mod]['input_img_path'] = s
        if mod == ref_mod:
            strip = s
    cfg.pop('image_shape_px', None)
    cfg = initialize_geometry_and_resolutions(cfg)          # bounds / shape of the strip; header only
    tiles = tiles_gdf.iloc[me.tile_idx].reset_index(drop=True)
    return ZoneShard(rank, world, cfg, tiles, strip, (me.in_r0, me.in_r1), (me.out_r0, me.out_r1), all_rows)


def run_zone_shard(model, shard: ZoneShard, patch_sizes: Dict[str, int]) -> Dict[str, torch.Tensor]:
    """This rank's strip through ``inference_and_write`` -> {task: uint8 (count, owned rows, W) device tensor}: the rows of
    the zone's class raster this rank owns (a view into the strip's result; nothing is written to disk).  {} for an empty
    shard."""
    if shard.raster is None:
        return {}
    ds = prep_dataset(shard.config, shard.tiles, patch_sizes)
    outs, _ = init_outputs(shard.config, shard.raster, 0)
    for sink in outs.values():
        sink.write_files = False                          # strips are assembled first; rank 0 writes the zone's files
    inference_and_write(model, ds, shard.tiles, shard.config, outs, shard.raster)
    o0, o1 = shard.out_rows[0] - shard.in_rows[0], shard.out_rows[1] - shard.in_rows[0]
    owned = {task: sink.device_array[:, o0:o1] for task, sink in outs.items()}
    for sink in outs.values():
        sink.release()
    return owned


def gather_row_strips(local: Optional[torch.Tensor], all_rows, full: Optional[torch.Tensor], rank: int, world: int,
                      dst: int = 0) -> None:
    """Rank ``dst`` receives every rank's owned rows into ``full`` (count, H, W) at ``all_rows[r]``; the others send
    theirs.  Point-to-point ``torch.distributed`` transfers of the result bytes (NCCL over NVLink in a GPU job, gloo on the
    CPU): 3.6 GB in total for a 60 000 x 60 000 zone, after the compute -- the data path itself has no collective."""
    import torch.distributed as dist
    if rank == dst:
        for r, (a, b) in enumerate(all_rows):
            if b <= a:
                continue
            if r == dst:
                full[:, a:b].copy_(local)
            else:
                buf = torch.empty((full.shape[0], b - a, full.shape[2]), dtype=full.dtype, device=full.device)
                dist.recv(buf, src=r)
                full[:, a:b].copy_(buf)
    elif local is not None and local.numel():
        dist.send(local.contiguous(), dst=dst)


def postpro_outputs(temp_paths: Dict[str, str], config: Dict) -> Dict[str, str]:
    """inference.py:633-641: with ``cog_conversion`` every output raster becomes ``<name>_COG.tif`` and the plain file is
    removed.  Returns the paths that exist afterwards (the reference returns nothing)."""
    final = dict(temp_paths)
    if config.get("cog_conversion", False):
        for task_name, temp_path in temp_paths.items():
            cog_path = temp_path.replace(".tif", "_COG.tif")
            convert_to_cog(temp_path, cog_path)
            final[task_name] = cog_path
            logger.info(f"\n[✓] Converted to COG: {cog_path}")
    return final


def run_inference(config_path: str) -> Dict[str, str]:
    """inference.py:644-674, repaired: config in -> rasters out.  Returns the
    written paths per task."""
    from torch.utils.data import DataLoader
    t0 = time.time()
    config = prep_config(config_path)
    ref_path = config['modalities'][config['reference_modality']]['input_img_path']
    tiles_gdf = generate_patches_from_reference(config, ref_path, None)
    logger.info(f"[✓] Sliced into {len(tiles_gdf)} tiles")
    patch_sizes = compute_patch_sizes(config)
    model = build_inference_model(config, patch_sizes).to(config['device'])
    ref_img = open_raster(ref_path)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        # one zone over the job's GPUs (the reference has no multi-GPU inference): row strips, rank 0 writes the files
        rank, world = dist.get_rank(), dist.get_world_size()
        shard = shard_zone(config, tiles_gdf, rank, world)
        owned = run_zone_shard(model, shard, patch_sizes)
        all_rows = shard.all_out_rows
        output_files = init_outputs(config, ref_img, 0)[0] if rank == 0 else {}
        for t in [t['name'] for t in config['tasks'] if t['active']]:
            gather_row_strips(owned.get(t), all_rows, output_files[t].device_array if rank == 0 else None, rank, world)
        if rank != 0:
            return {}
        torch.cuda.synchronize(config['device'])
        for dst in output_files.values():
            dst.close()
    else:
        dataset = prep_dataset(config, tiles_gdf, patch_sizes)
        dataloader = DataLoader(dataset, batch_size=config.get('batch_size', 8), num_workers=0)
        output_files, temp_paths = init_outputs(config, ref_img, 0)
        inference_and_write(model, dataloader, tiles_gdf, config, output_files, ref_img)
    written = postpro_outputs({k: v.written_path for k, v in output_files.items()}, config)       # inference.py:669
    logger.info(f"[✓] Total time: {time.time() - t0:.2f}s")
    return written

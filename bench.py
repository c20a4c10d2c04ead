#!/usr/bin/env python
"""bench.py -- zonal inference throughput (Mpx/s) of the B200 implementation, BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; ranks own row strips, no collective.)

Workload (BASELINE.json configs[1]): ConvNeXtV2-base + U-Net, synthetic 4-band uint8 raster
10000 x 10000 at 0.2 m/px, tile 512, margin 64 (overlap 128) -> 729 tiles, 19 classes, random-init
weights in the reference's checkpoint layout loaded through build_inference_model().
A step = one whole zone.  At N > 1 GPUs the workload is configs[3]: ONE 60000 x 60000 zone (24 649 tiles) whose tile rows
are dealt to the ranks as contiguous row strips (strong scaling, no data-path collective); value = zone pixels /
max-over-ranks time, per-rank times are reported.  The line also carries a "train" block: the configs[4] training step
(fwd + bwd + AdamW + DDP all-reduce), a few steps, samples/s.

One JSON line on stdout (rank 0).  `value`: raster resident in HBM, CUDA-event timed.
`e2e`: the same zone through inference_and_write() from pinned HOST memory to a HOST class raster
(H2D + D2H inside the timed region; TIFF encoding excluded).  `roofline`: all tcgen05 GEMM launches
of one batch, timed live with CUDA events on the launching stream.  `cpu_baseline`: the oracle
(torch fp32 eager restatement of the reference) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

PATCH, MARGIN, RES = 512, 64, 0.2
ARCH = "convnextv2_base-unet"
TASK = "AERIAL_LABEL-COSIA"
N_CLS = 19
LEFT, TOP = 700000.0, 6600000.0
GFLOP_PER_TILE = 189.72
# algorithmic GFLOP per 512^2 tile of the other architectures
ARCH_GFLOP = {"convnextv2_base-unet": 189.72, "swin_base_patch4_window12_384-upernet": 196.0, "resnet34-unet": 64.28}
METRIC = "zonal_inference_mpx_per_s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def zonal_config(weights_path: str, out_dir: str, raster_name: str, batch: int) -> dict:
    """The reference's zonal YAML schema (configs/config_model_zonal_segmentation.yaml) as a dict."""
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS
    return {
        "output_path": out_dir, "output_name": "bench_zone", "write_dataframe": False, "output_type": "argmax",
        "cog_conversion": False, "model_weights": weights_path, "use_gpu": True, "batch_size": batch,
        "num_worker": 0, "img_pixels_detection": PATCH, "margin": MARGIN, "output_px_meters": RES,
        "monotemp_arch": ARCH, "multitemp_model_ref_date": "05-15",
        "modalities": {
            "inputs": {"AERIAL_RGBI": True, "AERIAL-RLT_PAN": False, "DEM_ELEV": False, "SPOT_RGBI": False,
                       "SENTINEL2_TS": False, "SENTINEL1-ASC_TS": False, "SENTINEL1-DESC_TS": False},
            "AERIAL_RGBI": {"input_img_path": raster_name, "channels": [1, 2, 3, 4],
                            "normalization": {"type": "custom", "means": DEFAULT_MEANS, "stds": DEFAULT_STDS}},
        },
        "tasks": [{"name": TASK, "active": True, "class_names": {i: f"class_{i}" for i in range(N_CLS)}}],
    }


def random_state(mods: dict, seed: int = 2025, arch: str = ARCH) -> dict:
    """Seeded random state_dict in the reference's layout for the modalities ``{MOD: channels}`` (product model class only;
    the oracle is not involved)."""
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    from flair_for_aigle_b200.synthetic import randomize_state_
    c = zonal_config("unused", tempfile.gettempdir(), "unused", 1)
    c["monotemp_arch"] = arch
    for m, ch in mods.items():
        c["modalities"]["inputs"][m] = True
        c["modalities"].setdefault(m, {"input_img_path": "unused", "channels": list(range(1, ch + 1))})
    if "DEM_ELEV" in mods:
        c["modalities"]["DEM_ELEV"].update({"calc_elevation": True, "calc_elevation_stack_dsm": False})
    for m in list(c["modalities"]["inputs"]):
        if m not in mods:
            c["modalities"]["inputs"][m] = False
    sd = FLAIR_HUB_Model(prepare_model_config(c), {m: PATCH for m in mods}).state_dict()
    randomize_state_(sd, seed)
    return sd


def make_weights(path: str, seed: int = 2025) -> None:
    """Random-init checkpoint in the reference's state_dict layout (.safetensors)."""
    from safetensors.torch import save_file
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    from flair_for_aigle_b200.synthetic import randomize_state_
    cfg = prepare_model_config(zonal_config(path, tempfile.gettempdir(), "unused", 1))
    m = FLAIR_HUB_Model(cfg, {"AERIAL_RGBI": PATCH})
    sd = m.state_dict()
    randomize_state_(sd, seed)
    save_file({k: v.contiguous() for k, v in sd.items()}, path)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.rows = gpu_index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                          "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline_sample(weights_path: str, raster: np.ndarray, budget_s: float, max_tiles: int, batch: int = 4):
    """Oracle pipeline (reference CPU torch path) on a bounded sample of the zone's tiles."""
    from safetensors.torch import load_file
    from oracle.grid import Georef, generate_patches, tile_plan
    from oracle.models import FlairHubOracle
    from oracle.pipeline import load_batch
    from oracle.convert import write_tiles
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = FlairHubOracle(ARCH, {"AERIAL_RGBI": 4}, {TASK: N_CLS}).eval()
    model.load_state_dict(load_file(weights_path), strict=True)
    h, w = raster.shape[1:]
    geo = Georef(LEFT, TOP, RES, w, h)
    tiles = generate_patches(PATCH, MARGIN, RES, geo)
    plan = tile_plan(tiles, geo, PATCH, MARGIN)
    out = np.zeros((h, w), np.uint8)
    done, t_spent = 0, 0.0
    with torch.no_grad():
        # untimed warm-up (thread pools, oneDNN primitives)
        b = load_batch(raster, plan, [0], PATCH, DEFAULT_MEANS, DEFAULT_STDS, TASK, N_CLS)
        model(b)
        order = list(range(0, len(tiles), max(1, len(tiles) // max_tiles)))[:max_tiles]
        for s in range(0, len(order), batch):
            idx = order[s:s + batch]
            t0 = time.perf_counter()
            b = load_batch(raster, plan, idx, PATCH, DEFAULT_MEANS, DEFAULT_STDS, TASK, N_CLS)
            logits, _ = model(b)
            write_tiles(logits[TASK].numpy(), plan[idx], MARGIN, out, "argmax")
            t_spent += time.perf_counter() - t0
            done += len(idx)
            if t_spent > budget_s:
                break
    s_per_tile = t_spent / done
    mpx_s = (w * h / 1e6) / (s_per_tile * len(tiles))
    return {"value": round(mpx_s, 4), "unit": "Mpx/s", "cores": threads, "kind": "port",
            "sample": f"{done} of {len(tiles)} tiles of the {w}x{h} zone through the oracle pipeline "
                      f"(read+normalise+forward+crop/argmax+write), {s_per_tile*1e3:.0f} ms/tile, extrapolated by tile count; "
                      f"torch {torch.__version__} fp32 eager, {threads} threads"}, s_per_tile


def zone_side(args, world: int) -> int:
    """N = 1: BASELINE.json configs[1] (10 000 x 10 000, the configuration the metric is quoted on).
    N > 1: configs[3], ONE 60 000 x 60 000 zone (24 649 tiles, 157 tile rows) cut into N row strips -- strong scaling.
    ``--zone`` / FZ_BENCH_ZONE override (e.g. the 60 k zone on one GPU, the strong-scaling base, profiles/)."""
    if args.zone:
        return args.zone
    if os.environ.get("FZ_BENCH_ZONE"):
        return int(os.environ["FZ_BENCH_ZONE"])
    return 10000 if world == 1 else 60000


def run_reference(args, rank: int, world: int) -> None:
    """The reference's CPU path (the oracle pipeline: the reference cannot be installed here, smp / timm / rasterio are
    absent) on the host cores, for our arm's zone.  Each step = a fixed sample of 32 tiles, extrapolated by tile count."""
    if rank != 0:
        return
    from flair_for_aigle_b200.synthetic import synthetic_raster
    from oracle.grid import Georef, generate_patches
    side = zone_side(args, world)
    tmp = tempfile.mkdtemp(prefix="fz_bench_ref_")
    wpath = os.path.join(tmp, "weights.safetensors")
    make_weights(wpath)
    # a 10000 x 2048 band holds 162 tiles: plenty for the sample; same tile size, margin, model = same per-tile cost
    raster = synthetic_raster(10000, 10000, row0=0, rows=2048)
    per_tile, cb = [], None
    for i in range(args.steps + args.warmup):
        cb, spt = cpu_baseline_sample(wpath, raster, budget_s=1e9, max_tiles=32)
        if i >= args.warmup:
            per_tile.append(spt)
    spt = float(np.mean(per_tile))
    n_tiles = len(generate_patches(PATCH, MARGIN, RES, Georef(LEFT, TOP, RES, side, side)))
    mpx_s = (side * side / 1e6) / (spt * n_tiles)
    cb["value"] = round(mpx_s, 4)
    cb["sample"] = (f"32 tiles per step, {args.steps} timed steps (spread {min(per_tile) * 1e3:.0f}-{max(per_tile) * 1e3:.0f} ms/tile), "
                    + cb["sample"])
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mpx_s, 4), "unit": "Mpx/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(spt * n_tiles * 1e3, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(side, n_tiles, world)},
        "cpu_baseline": cb,
        "e2e": {"value": round(mpx_s, 4), "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_string(side: int, n_tiles: int, world: int) -> str:
    return (f"{ARCH} zonal inference, synthetic {side}x{side}x4 uint8 @0.2m, tile {PATCH} margin {MARGIN} "
            f"(overlap {2 * MARGIN}), {n_tiles} tiles, {N_CLS} classes, argmax raster"
            + ("" if world == 1 else f"; ONE zone cut into {world} row strips (tile rows dealt contiguously, margin halo re-read)"))


def other_config_block(dev, arch: str, side: int, steps: int, batch: int, label: str):
    """One more BASELINE.json configuration on this GPU, device-timed like `value` (raster resident in HBM, CUDA graph, CUDA
    events, inputs larger than L2): e.g. configs[2], Swin-base + UPerNet on a 20 000 x 20 000 zone."""
    global ARCH
    from flair_for_aigle_b200.engine.zonal import ZonalRunner
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model
    from flair_for_aigle_b200.flair_zonal_detection.raster import ZoneRaster, register_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import (generate_patches_from_reference, ownership_windows,
                                                                    tile_plan)
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS, synthetic_raster_to_pinned
    saved, ARCH = ARCH, arch
    try:
        tmp = tempfile.mkdtemp(prefix="fz_bench_other_")
        wpath = os.path.join(tmp, "weights.safetensors")
        make_weights(wpath)
        name = f"synthetic://other_{side}"
        register_raster(name, ZoneRaster(np.broadcast_to(np.zeros((1, 1, 1), np.uint8), (4, side, side)), LEFT, TOP, RES))
        cfg = inf.initialize_geometry_and_resolutions(zonal_config(wpath, tmp, name, batch))
        cfg["device"] = dev
        tiles = generate_patches_from_reference(cfg, name, None)
        plan = tile_plan(tiles, cfg["image_bounds"], RES, PATCH, MARGIN)
        own = ownership_windows(plan)
        model = build_inference_model(cfg, {"AERIAL_RGBI": PATCH}).to(dev)
        runner = ZonalRunner(model.engine(TASK, max_batch=batch), MARGIN, use_graph=True, norm=(DEFAULT_MEANS, DEFAULT_STDS))
        host = torch.empty((4, side, side), dtype=torch.uint8, pin_memory=True)
        synthetic_raster_to_pinned(side, side, host, dev)
        raster_dev, _ = runner.buffers((4, side, side), (side, side), torch.uint8)
        raster_dev.copy_(host, non_blocking=True)
        out = torch.zeros((side, side), dtype=torch.uint8, device=dev)
        for _ in range(2):
            runner.run(raster_dev, plan, own, out)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            runner.run(raster_dev, plan, own, out)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        pk, _ = peaks()
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        res = {"config": label, "value": round(side * side / 1e6 / (ms / 1e3), 2), "unit": "Mpx/s", "ms_per_step": round(ms, 2),
               "steps": steps, "warmup": 2, "tiles": len(tiles), "tiles_per_s": round(len(tiles) / (ms / 1e3), 1),
               "model_flops_frac_of_peak": round(ARCH_GFLOP[arch] * 1e9 * len(tiles) / (ms * 1e-3) / 1e12 / peak, 4),
               "class_raster_checksum": int(out.to(torch.int64).sum().item())}
        del runner, raster_dev, out, model, host
        torch.cuda.empty_cache()
        return res
    finally:
        ARCH = saved


def train_block(dev, rank: int, world: int, steps: int, warmup: int):
    """BASELINE.json configs[4]: convnextv2_base-unet, AERIAL_RGBI (16,4,512,512) + DEM_ELEV (16,1,512,512) per GPU, weighted
    CE (classes 15-18 weight 0), AdamW(5e-5, wd 0.01), DDP gradient all-reduce (tasks_module.py:133-167,377-391;
    trainers.py:81-91).  -> dict for the JSON line's "train" key (rank 0), None elsewhere."""
    import torch.distributed as dist
    from flair_for_aigle_b200.engine.convnext_unet import CONVNEXTV2_CFGS
    from flair_for_aigle_b200.engine.train_step import ConvNeXtUNetTrainer
    B, P = 16, PATCH
    mods = {"AERIAL_RGBI": 4, "DEM_ELEV": 1}
    state = {k: v.to(dev) for k, v in random_state(mods, seed=2025, arch="convnextv2_base-unet").items()}
    depths, dims = CONVNEXTV2_CFGS["convnextv2_base"]
    w = torch.ones(N_CLS, device=dev)
    w[15:] = 0
    # the step replayed as a CUDA graph (one process) or a chain of graphs cut at the gradient buckets (torch.distributed)
    graphed = os.environ.get("FZ_TRAIN_GRAPH", "1") != "0"
    tr = ConvNeXtUNetTrainer(state, depths, dims, list(mods), TASK, w, cuda_graph=graphed)
    g = torch.Generator(device="cpu").manual_seed(2025 + rank)
    host = {k: torch.randn(B, c, P, P, generator=g).pin_memory() for k, c in mods.items()}
    host[TASK] = torch.randint(0, N_CLS, (B, P, P), generator=g, dtype=torch.int32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host.values())

    def one_step():
        batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}     # H2D of the step's inputs
        loss, _ = tr.step(batch)
        return float(loss)                                                     # D2H of the loss

    losses = [one_step() for _ in range(max(2, warmup))]        # step 1 eager, step 2 captures the graph, then replays
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ar = []
    e0.record()
    for _ in range(steps):
        losses.append(one_step())
        ar.append(getattr(tr, "last_allreduce_ms", 0.0))
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / steps, sum(ar) / max(len(ar), 1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step, ms_ar = float(ms[0]), float(ms[1])
    drift = 0.0
    if world > 1:       # DDP invariant: after identical averaged updates every rank holds the same parameters
        ref = tr.opt.arena.clone()
        dist.broadcast(ref, 0)
        d = (tr.opt.arena - ref).abs().max().reshape(1)
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        drift = float(d)
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    del tr, state
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    pk, how = peaks()
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    tflop_step = 1.063 * B                       # 177.2 GMAC fwd / sample, x3 fwd+bwd = 1.063 TFLOP / sample
    return {"config": "configs[4]: convnextv2_base-unet, AERIAL_RGBI 4ch + DEM_ELEV 1ch, batch 16 x 512^2 per GPU, weighted CE, "
                      "AdamW, DDP all-reduce of the gradient arena; inputs from pinned host memory every step, loss read back",
            "samples_per_s": round(B * world / ms_step * 1e3, 2), "samples_per_s_per_gpu": round(B / ms_step * 1e3, 2),
            "ms_per_step": round(ms_step, 2), "steps": steps, "warmup": max(2, warmup), "cuda_graph": graphed,
            "allreduce_ms_exposed": round(ms_ar, 3), "allreduce": "bucketed per backward group (decoder, fusion, encoder stages deepest "
            "first), each bucket's NCCL all-reduce started on a side stream as its gradients land; exposed = what the compute "
            "stream waited at the end", "max_parameter_difference_across_ranks": drift, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
            "loss_first_last": [round(losses[0], 4), round(losses[-1], 4)], "peak_memory_gib": round(peak_mem, 1),
            "roofline": {"bound": "tensor", "achieved": round(tflop_step / (ms_step * 1e-3), 1), "peak": peak,
                         "unit": "TFLOP/s", "frac": round(tflop_step / (ms_step * 1e-3) / peak, 4),
                         "note": f"1.063 TFLOP per sample (fwd + bwd, SURVEY 8d) x {B} / step time; peak = {how} sustained bf16"},
            "dtype": "fp16 forward operands, bf16 gradient operands, fp32 accumulate / master weights"}


def file_io_block(inf, model, cfg, tiles, patch_sizes, host, tmp, dev, checksum, total_px) -> dict:
    """SURVEY 8(f) rank 1, measured: the same zone FILE TO FILE through the public API -- a tiled LZW GeoTIFF of the RGBI zone on
    disk (page cache warm) -> open_raster (libfz_rasterio.so decodes it block-parallel into page-locked memory) ->
    inference_and_write -> the class raster as a tiled LZW GeoTIFF (the reference's profile) -> optional COG conversion
    (postprocess.py:33-52).  Host work is on this box's cores; the reference does the same with one rasterio window read per
    tile and one LZW window write per tile on one core."""
    import torch
    from flair_for_aigle_b200 import raster_io
    from flair_for_aigle_b200.flair_zonal_detection.postprocess import convert_to_cog
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink, open_raster
    src = os.path.join(tmp, "zone_rgbi.tif")
    t0 = time.perf_counter()
    raster_io.write_geotiff(src, host.numpy(), LEFT, TOP, RES, epsg=2154, pixel_interleave=True, predictor=2)
    make_ms = (time.perf_counter() - t0) * 1e3
    cfg_f = dict(cfg)
    cfg_f["modalities"] = json.loads(json.dumps(cfg["modalities"]))
    cfg_f["modalities"]["AERIAL_RGBI"]["input_img_path"] = src
    cfg_f.pop("image_shape_px", None)
    out_dir = os.path.join(tmp, "file_io")
    os.makedirs(out_dir, exist_ok=True)
    cfg_f["output_path"] = out_dir
    cfg_f = inf.initialize_geometry_and_resolutions(cfg_f)      # header only: no pixel is decoded for the geometry
    cfg_f["device"] = dev
    RasterSink.write_files = True
    runs = []
    written = None
    for _ in range(2):                                          # the second run is the reported one
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        ref = open_raster(src)
        ds = inf.prep_dataset(cfg_f, tiles, patch_sizes)
        outs, _ = inf.init_outputs(cfg_f, ref, 0)
        # decode (background thread, bottom-up slabs) || H2D || forward || D2H, then the GeoTIFF written by close()
        inf.inference_and_write(model, ds, tiles, cfg_f, outs, ref)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        written = outs[TASK].written_path
        outs[TASK].release()
        runs.append(t1 - t0)
        del ds, ref, outs
    zone_s = runs[-1]
    t0 = time.perf_counter()
    raster_io.read_raster(src, out=host.numpy())                # the decode by itself, into the same pinned buffer
    decode_s = time.perf_counter() - t0
    got, info = raster_io.read_raster(written)
    same = bool(int(got.astype(np.int64).sum()) == checksum)
    t0 = time.perf_counter()
    raster_io.write_geotiff(os.path.join(out_dir, "again.tif"), got, LEFT, TOP, RES, epsg=2154)
    write_ms = (time.perf_counter() - t0) * 1e3
    out_mb = os.path.getsize(written) / 1e6
    t0 = time.perf_counter()
    cog = written.replace(".tif", "_COG.tif")
    convert_to_cog(written, cog)
    cog_ms = (time.perf_counter() - t0) * 1e3
    input_mb, cog_overviews = os.path.getsize(src) / 1e6, raster_io.tiff_info(cog).overviews
    shutil.rmtree(out_dir, ignore_errors=True)                  # ~350 MB of scratch files: gone before the next block
    try:
        os.remove(src)
    except OSError:
        pass
    return {"value": round(total_px / 1e6 / zone_s, 2), "unit": "Mpx/s",
            "ms": {"file_to_file": round(zone_s * 1e3, 1), "decode_input_geotiff_to_pinned_alone": round(decode_s * 1e3, 1),
                   "encode_class_geotiff_alone": round(write_ms, 1), "convert_to_cog_extra": round(cog_ms, 1),
                   "make_input_file_setup": round(make_ms, 1)},
            "input_file_mb": round(input_mb, 1), "output_file_mb": round(out_mb, 2),
            "output": {"tiled": info.tiled, "block": info.block_w, "compression": "lzw", "cog_overviews": cog_overviews},
            "same_result_as_value_leg": same, "host_cores": os.cpu_count(),
            "note": "GeoTIFF on disk -> open_raster -> inference_and_write -> LZW GeoTIFF on disk, all through the public API; "
                    "file codecs = libfz_rasterio.so, one 512x512 block per task on all host cores (profiles/r2_raster_io_bench.txt "
                    "has the libtiff comparison); the input decodes bottom-up on a background thread while run_streamed uploads and "
                    "computes the rows already there; value = zone px / wall time of open_raster + prep_dataset + init_outputs + "
                    "inference_and_write; COG conversion reported beside it, not inside"}


def main() -> None:
    global ARCH, GFLOP_PER_TILE
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("FZ_BENCH_BATCH", "37")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the configs[4] training-step block")
    ap.add_argument("--no-file-io", action="store_true", help="skip the GeoTIFF-in / GeoTIFF-out block")
    ap.add_argument("--arch", default=ARCH, choices=sorted(ARCH_GFLOP),
                    help="default = BASELINE.json's metric configuration; the others are measured for DESIGN.md only")
    ap.add_argument("--zone", type=int, default=0, help="zone side in pixels (default: 10000 at 1 GPU, 60000 sharded at N > 1)")
    args = ap.parse_args()
    ARCH, GFLOP_PER_TILE = args.arch, ARCH_GFLOP[args.arch]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep "NCCL version ..." off stdout: the JSON line stays alone there
        dist.init_process_group("nccl", device_id=dev)

    from flair_for_aigle_b200 import native as nv
    from flair_for_aigle_b200.engine.strips import shard_rows
    from flair_for_aigle_b200.engine.zonal import ZonalRunner
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink, ZoneRaster, register_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import (generate_patches_from_reference,
                                                                    ownership_windows, tile_plan)
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS, synthetic_raster, synthetic_raster_to_pinned

    tmp = tempfile.mkdtemp(prefix=f"fz_bench_{rank}_")
    wpath = os.path.join(tmp, "weights.safetensors")
    make_weights(wpath)

    # ---- ONE global zone; rank r owns row strip r of the global tile plan
    side = zone_side(args, world)
    gh = gw = side
    shape_only = ZoneRaster(np.broadcast_to(np.zeros((1, 1, 1), np.uint8), (4, gh, gw)), LEFT, TOP, RES)
    register_raster("synthetic://zone_shape", shape_only)
    cfg = zonal_config(wpath, tmp, "synthetic://zone_shape", args.batch)
    cfg = inf.initialize_geometry_and_resolutions(cfg)
    cfg["device"] = dev
    cfg["labels"] = [TASK]
    tiles = generate_patches_from_reference(cfg, "synthetic://zone_shape", None)
    gplan = tile_plan(tiles, cfg["image_bounds"], RES, PATCH, MARGIN)
    gown = ownership_windows(gplan)
    shard = shard_rows(gplan, gown, PATCH, gh, world)[rank]
    n_tiles_rank = len(shard.tile_idx)
    in_rows = shard.in_r1 - shard.in_r0
    out_rows = shard.out_r1 - shard.out_r0
    log(f"[rank {rank}] zone {gw}x{gh}: {len(tiles)} tiles, this rank {n_tiles_rank} tiles, input rows "
        f"[{shard.in_r0},{shard.in_r1}), output rows [{shard.out_r0},{shard.out_r1})")

    # ---- this rank's input strip in pinned host memory (generated on the GPU in row chunks: the numpy generator makes
    #      ~10 MB/s and the 60k zone is 14.4 GB; set-up, outside every timed region)
    t_setup = time.perf_counter()
    host = torch.empty((4, in_rows, gw), dtype=torch.uint8, pin_memory=True)
    synthetic_raster_to_pinned(gh, gw, host, dev, row0=shard.in_r0)
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    log(f"[rank {rank}] raster strip {host.numel() / 1e9:.2f} GB generated + pinned in {time.perf_counter() - t_setup:.1f} s")
    strip = ZoneRaster.from_pinned(host, LEFT, TOP - shard.in_r0 * RES, RES, name="synthetic://strip")
    register_raster("synthetic://strip", strip)

    patch_sizes = {"AERIAL_RGBI": PATCH}
    model = build_inference_model(cfg, patch_sizes).to(dev)
    eng = model.engine(TASK, max_batch=args.batch)
    runner = ZonalRunner(eng, MARGIN, use_graph=True, norm=(DEFAULT_MEANS, DEFAULT_STDS))

    # the raster is uploaded straight into the buffer the runner's CUDA graph reads (no second device copy of a 14 GB strip)
    raster_dev, _ = runner.buffers((4, in_rows, gw), (out_rows, gw), torch.uint8)
    raster_dev.copy_(host, non_blocking=True)
    out_dev = torch.zeros((out_rows, gw), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- value: inputs resident in HBM
    for _ in range(args.warmup):
        runner.run(raster_dev, shard.plan, shard.own, out_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nb = 0
    for _ in range(args.steps):
        nb += runner.run(raster_dev, shard.plan, shard.own, out_dev)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_rank = e0.elapsed_time(e1) / args.steps
    rank_ms = [ms_rank]
    rank_tiles = [n_tiles_rank]
    if use_dist:
        t = torch.tensor([ms_rank, float(n_tiles_rank)], device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_ms = [float(x[0]) for x in allt]
        rank_tiles = [int(x[1]) for x in allt]
    ms_per_step = max(rank_ms)
    total_px = gw * gh
    value = total_px / 1e6 / (ms_per_step / 1e3)
    launches = nb * runner.count_launches()
    checksum = int(out_dev.to(torch.int64).sum().item())

    # ---------------------------------------------------------------- e2e: public API on this rank's shard of the GLOBAL plan
    RasterSink.write_files = False
    cfg_e = dict(cfg)
    cfg_e["modalities"] = json.loads(json.dumps(cfg["modalities"]))
    cfg_e["modalities"]["AERIAL_RGBI"]["input_img_path"] = "synthetic://strip"
    cfg_e = inf.initialize_geometry_and_resolutions(cfg_e)   # raster = this rank's strip (its input rows of the zone)
    cfg_e["device"] = dev
    # the rank's rows of the GLOBAL tile table (zone coordinates): the strip raster carries the zone's georeferencing, so
    # the same tiles address the same pixels; a tile's ownership inside the shard equals its global ownership on the
    # rows this rank owns (a global last writer that is in the shard is also the shard's last writer)
    tiles_e = tiles.iloc[shard.tile_idx].reset_index(drop=True)
    o0, o1 = shard.out_r0 - shard.in_r0, shard.out_r1 - shard.in_r0

    def e2e_step():
        ds = inf.prep_dataset(cfg_e, tiles_e, patch_sizes)           # fresh dataset: raster is uploaded again
        outs, _ = inf.init_outputs(cfg_e, strip, 0)
        inf.inference_and_write(model, ds, tiles_e, cfg_e, outs, strip)   # H2D + compute + D2H (close())
        res = outs[TASK].to_host()
        nbytes = res.size
        chk = int(res[0, o0:o1].astype(np.int64).sum()) if e2e_step.check else None
        outs[TASK].release()                                         # recycle the pinned result buffer
        return nbytes, chk

    e2e_step.check = True
    _, chk = e2e_step()
    e2e_same = bool(chk == checksum)
    e2e_step.check = False
    for _ in range(max(0, min(args.warmup, 2) - 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_bytes, _ = e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / args.steps
    if use_dist:
        t = torch.tensor([e2e_s, float(e2e_same)], device=dev)
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        e2e_s, e2e_same = float(tmax[0]), bool(tmin[1] > 0.5)
        hb = torch.tensor([float(4 * in_rows * gw), float(res_bytes)], device=dev)
        dist.all_reduce(hb)
        h2d_total, d2h_total = int(hb[0]), int(hb[1])
    else:
        h2d_total, d2h_total = int(4 * in_rows * gw), int(res_bytes)
    e2e_val = total_px / 1e6 / e2e_s

    # ---------------------------------------------------------------- file to file, after the contract legs
    file_io = None
    if rank == 0 and world == 1 and side <= 20000 and not args.no_file_io:
        try:
            file_io = file_io_block(inf, model, cfg, tiles, patch_sizes, host, tmp, dev, checksum, total_px)
        except Exception as ex:  # noqa: BLE001 -- the contract line must survive a failure of the extra block
            file_io = {"error": repr(ex)}
            log(f"[rank {rank}] file I/O block failed: {ex!r}")
        RasterSink.write_files = False

    # ---------------------------------------------------------------- roofline: GEMM launches of one batch, live
    roof = None
    breakdown = {}
    if rank == 0:
        runner_e = ZonalRunner(eng, MARGIN, use_graph=False, norm=(DEFAULT_MEANS, DEFAULT_STDS))
        sub_plan, sub_own = shard.plan[:args.batch], shard.own[:args.batch]
        runner_e.run(raster_dev, sub_plan, sub_own, out_dev)
        torch.cuda.synchronize(dev)
        nv.PROFILE = []
        for _ in range(3):
            runner_e.run(raster_dev, sub_plan, sub_own, out_dev)
        torch.cuda.synchronize(dev)
        prof, nv.PROFILE = nv.PROFILE, None
        fam_ms, tc = {}, {"gemm_tcgen05": [0.0, 0.0, 0], "conv3x3_tcgen05": [0.0, 0.0, 0]}
        for fam, meta, a, b in prof:
            d = a.elapsed_time(b)
            fam_ms[fam] = fam_ms.get(fam, 0.0) + d
            if fam == "gemm_tcgen05":
                fl = 2.0 * meta["M"] * meta["N"] * meta["K"]
            elif fam == "conv3x3_tcgen05":
                fl = 2.0 * meta["B"] * meta["H"] * meta["H"] * 9 * meta["Cin"] * meta["Cout"]
            else:
                continue
            tc[fam][0] += fl
            tc[fam][1] += d
            tc[fam][2] += 1
        dom = max(tc, key=lambda k: tc[k][1])            # the tensor-core family with the largest time share
        gemm_flops, gemm_ms, gemm_n = tc[dom]
        tot = sum(fam_ms.values())
        breakdown = {k: round(v / tot, 4) for k, v in sorted(fam_ms.items(), key=lambda kv: -kv[1])}
        pk, how = peaks()
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12
        traffic = None
        try:        # dram read+write per launch of the dominant GEMM shape, from the committed ncu --set full capture
            if ARCH == "convnextv2_base-unet":
                with open(os.path.join(ROOT, "profiles", "r2_ncu_gemm_traffic.json")) as f:
                    tj = json.load(f)
                traffic = {"bytes_per_launch": tj["dram_bytes_per_launch"],
                           "algorithmic_bytes_per_launch": tj["algorithmic_bytes_per_launch"], "kernel": tj["kernel"],
                           "source": tj["source"]}
        except Exception:
            traffic = None
        per_batch = gemm_n // 3
        roof = {"bound": "tensor", "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 4), "traffic": traffic,
                "kernel": (f"gemm_bf16_kernel / gemm_bf16_pair_kernel (tcgen05, all {per_batch} linear / 1x1 GEMM launches of one "
                           f"{ARCH} batch of {args.batch} tiles, fp16 operands)"
                           if dom == "gemm_tcgen05" else
                           f"conv3x3_kernel / conv3x3_rows_kernel (tcgen05 implicit GEMM, all {per_batch} 3x3 convolutions of one {ARCH} batch)"),
                "peak_source": f"{how} bf16_tflops_sustained (kernel timed inside a long step; kind::f16 runs fp16 and bf16 at "
                               "the same rate)",
                "launches_timed": gemm_n, "launches_per_batch": per_batch, "avg_launch_us": round(gemm_ms / gemm_n * 1e3, 2),
                "share_of_step_eager": breakdown.get(dom),
                "model_flops_frac_of_peak": round(GFLOP_PER_TILE * 1e9 * max(rank_tiles) / (ms_per_step * 1e-3) / 1e12 / peak, 4)}

    # free the zone before the training block (the 60k strip and its graph buffers are tens of GB)
    del runner, raster_dev, out_dev
    if rank == 0:
        del runner_e
    torch.cuda.empty_cache()

    others = None
    if rank == 0 and world == 1 and not args.no_train and ARCH == "convnextv2_base-unet" and side == 10000:
        try:    # BASELINE.json configs[2]: Swin-base + UPerNet, 20 000 x 20 000 zone (2809 tiles), device-timed
            others = [other_config_block(dev, "swin_base_patch4_window12_384-upernet", 20000, 2, args.batch,
                                         "configs[2]: swin_base_patch4_window12_384-upernet zonal inference, synthetic "
                                         "20000x20000x4 uint8 @0.2m, tile 512 margin 64, 19 classes, argmax raster")]
        except Exception as ex:  # noqa: BLE001
            others = [{"error": repr(ex)}]
            log(f"[rank {rank}] configs[2] block failed: {ex!r}")

    train = None
    if not args.no_train and ARCH == "convnextv2_base-unet":
        try:
            train = train_block(dev, rank, world, steps=min(args.steps, 5), warmup=2)
        except Exception as ex:  # noqa: BLE001 -- the inference line must survive a training failure
            train = {"error": repr(ex)} if rank == 0 else None
            log(f"[rank {rank}] training block failed: {ex!r}")

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            band = synthetic_raster(10000, 10000, row0=0, rows=2048)
            cpu_base, _ = cpu_baseline_sample(wpath, band, budget_s=20.0, max_tiles=32)
        except Exception as ex:  # noqa: BLE001
            cpu_base = {"value": None, "unit": "Mpx/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"failed: {ex!r}"}

    if rank == 0:
        strong = None
        if world > 1:
            strong = {"zone": f"{gw}x{gh}", "tiles": len(tiles), "tile_rows": int(len(np.unique(gplan[:, 2]))),
                      "tiles_per_rank": rank_tiles, "ms_per_rank": [round(v, 2) for v in rank_ms],
                      "imbalance_max_over_mean": round(max(rank_ms) / (sum(rank_ms) / len(rank_ms)), 4),
                      "tiles_per_s_per_gpu": round(len(tiles) / world / (ms_per_step / 1e3), 1),
                      "halo_rows_reread_per_strip": int(in_rows - out_rows),
                      "note": "no data-path collective: every rank reads its rows (+ margin halo) of the zone and owns the "
                              "output rows of its tile rows; 157 tile rows do not divide by 4 or 8, which is the imbalance; "
                              "the 1-GPU value of THIS zone (strong-scaling base) is in profiles/r2_bench_60k_1gpu.json"}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mpx/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16" if nv.op_dtype() == torch.float16 else "bf16", "data": "synthetic",
            "config": {
                "workload": workload_string(side, len(tiles), world),
                "batch_tiles": args.batch, "tiles_per_gpu": max(rank_tiles), "cuda_graph": True,
                "operands": "fp16 operands, fp32 accumulate (tcgen05 kind::f16), fp32 residual stream / statistics",
                "l2": f"inputs larger than L2: {4 * in_rows * gw / 1e6:.0f} MB raster strip, >126 MB of activations per batch",
                "tiles_per_s": round(len(tiles) / (ms_per_step / 1e3), 1), "class_raster_checksum": checksum},
            "e2e": {"value": round(e2e_val, 2), "unit": "Mpx/s", "h2d_bytes_per_step": h2d_total,
                    "d2h_bytes_per_step": d2h_total, "same_result_as_value_leg": e2e_same,
                    "note": "inference_and_write() on each rank's rows of the global tile table: pinned host raster strip -> "
                            "HBM, fused forward, class raster -> pinned host (bytes summed over ranks); file encoding excluded"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "kernel_time_shares_eager": breakdown,
            "strong_scaling": strong,
            "other_configs": others,
            "train": train,
            "file_io": file_io,
        }
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

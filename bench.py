#!/usr/bin/env python
"""bench.py -- zonal inference throughput (Mpx/s) of the B200 implementation, BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; ranks own row strips, no collective.)

Workload (BASELINE.json configs[1]): ConvNeXtV2-base + U-Net, synthetic 4-band uint8 raster
10000 x 10000 at 0.2 m/px, tile 512, margin 64 (overlap 128) -> 729 tiles, 19 classes, random-init
weights in the reference's checkpoint layout loaded through build_inference_model().
A step = one whole zone.  At N GPUs the zone is 10000 x (10000*N) and each rank owns one
10000-row strip (weak scaling); value = all pixels / max-over-ranks time.

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
ZONE_W, ZONE_H = 10000, 10000
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


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    from flair_for_aigle_b200.synthetic import synthetic_raster
    tmp = tempfile.mkdtemp(prefix="fz_bench_ref_")
    wpath = os.path.join(tmp, "weights.safetensors")
    make_weights(wpath)
    sample_h = 2048   # a 10000 x 2048 band of the zone holds plenty of tiles for the sample
    raster = synthetic_raster(ZONE_H, ZONE_W, row0=0, rows=sample_h)
    # tiles of the band are the zone's tiles (grid is bottom-anchored per raster; use the band as its own zone
    # for the sample -- same tile size, margin, model: per-tile cost is identical)
    vals, per_tile = [], []
    total = args.steps + args.warmup
    for i in range(total):
        cb, spt = cpu_baseline_sample(wpath, raster, budget_s=max(4.0, 40.0 / total), max_tiles=8)
        if i >= args.warmup:
            per_tile.append(spt)
    spt = float(np.mean(per_tile))
    # the N-GPU arm works on a 10000 x (10000*N) zone (weak scaling): same zone here, same per-tile cost
    from oracle.grid import Georef, generate_patches
    zone_h = ZONE_H * max(1, world)
    n_tiles = len(generate_patches(PATCH, MARGIN, RES, Georef(LEFT, TOP, RES, ZONE_W, zone_h)))
    mpx_s = (ZONE_W * zone_h / 1e6) / (spt * n_tiles)
    cb["value"] = round(mpx_s, 4)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mpx_s, 4), "unit": "Mpx/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(spt * n_tiles * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{ARCH} zonal inference, synthetic {ZONE_W}x{zone_h}x4 uint8 @0.2m, tile {PATCH} "
                               f"margin {MARGIN} ({n_tiles} tiles), {N_CLS} classes; CPU oracle (the reference cannot be "
                               "imported: smp/timm/rasterio absent), each step = bounded tile sample extrapolated"},
        "cpu_baseline": cb,
        "e2e": {"value": round(mpx_s, 4), "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    global ARCH, GFLOP_PER_TILE, ZONE_W, ZONE_H
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("FZ_BENCH_BATCH", "37")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--arch", default=ARCH, choices=sorted(ARCH_GFLOP),
                    help="default = BASELINE.json's metric configuration; the others are measured for DESIGN.md only")
    ap.add_argument("--zone", type=int, default=ZONE_W, help="zone side in pixels per GPU (default 10000)")
    args = ap.parse_args()
    ARCH, GFLOP_PER_TILE, ZONE_W, ZONE_H = args.arch, ARCH_GFLOP[args.arch], args.zone, args.zone

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
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, compute_patch_sizes
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink, ZoneRaster, register_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import (generate_patches_from_reference,
                                                                    ownership_windows, tile_plan)
    from flair_for_aigle_b200.synthetic import synthetic_raster

    tmp = tempfile.mkdtemp(prefix=f"fz_bench_{rank}_")
    wpath = os.path.join(tmp, "weights.safetensors")
    make_weights(wpath)

    # ---- global zone: 10000 x (10000 * world); rank r owns row strip r
    gh, gw = ZONE_H * world, ZONE_W
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

    # ---- this rank's input strip, generated straight into pinned host memory
    host = torch.empty((4, in_rows, gw), dtype=torch.uint8, pin_memory=True)
    synthetic_raster(gh, gw, row0=shard.in_r0, rows=in_rows, out=host.numpy())
    strip = ZoneRaster.from_pinned(host, LEFT, TOP - shard.in_r0 * RES, RES, name="synthetic://strip")
    register_raster("synthetic://strip", strip)

    patch_sizes = {"AERIAL_RGBI": PATCH}
    model = build_inference_model(cfg, patch_sizes).to(dev)
    eng = model.engine(TASK, max_batch=args.batch)
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS
    runner = ZonalRunner(eng, MARGIN, use_graph=True, norm=(DEFAULT_MEANS, DEFAULT_STDS))

    raster_dev = host.to(dev, non_blocking=True)
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
    ms = e0.elapsed_time(e1)
    if use_dist:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    total_px = gw * gh
    value = total_px / 1e6 / (ms_per_step / 1e3)
    launches = nb * runner.count_launches()
    checksum = int(out_dev.to(torch.int64).sum().item())

    # ---------------------------------------------------------------- e2e: public API, host buffers
    RasterSink.write_files = False
    cfg_e = dict(cfg)
    cfg_e["modalities"] = json.loads(json.dumps(cfg["modalities"]))
    cfg_e["modalities"]["AERIAL_RGBI"]["input_img_path"] = "synthetic://strip"
    cfg_e = inf.initialize_geometry_and_resolutions(cfg_e)   # the strip as its own zone (this rank's work)
    cfg_e["device"] = dev
    tiles_e = generate_patches_from_reference(cfg_e, "synthetic://strip", None)

    def e2e_step():
        ds = inf.prep_dataset(cfg_e, tiles_e, patch_sizes)           # fresh dataset: raster is uploaded again
        outs, _ = inf.init_outputs(cfg_e, strip, 0)
        inf.inference_and_write(model, ds, tiles_e, cfg_e, outs, strip)   # H2D + compute + D2H (close())
        res = outs[TASK].to_host()
        nbytes = res.size
        outs[TASK].release()                                         # recycle the pinned result buffer
        return nbytes

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_bytes = e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / args.steps
    if use_dist:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_px = gw * in_rows * world
    e2e_val = e2e_px / 1e6 / e2e_s

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
                with open(os.path.join(ROOT, "profiles", "r1_ncu_gemm_traffic.json")) as f:
                    tj = json.load(f)
                traffic = {"bytes_per_launch": tj["dram_bytes_per_launch"],
                           "algorithmic_bytes_per_launch": tj["algorithmic_bytes_per_launch"], "kernel": tj["kernel"],
                           "source": tj["source"]}
        except Exception:
            traffic = None
        roof = {"bound": "tensor", "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 4), "traffic": traffic,
                "kernel": (f"gemm_bf16_kernel / gemm_bf16_pair_kernel (tcgen05, every linear / 1x1 GEMM of one {ARCH} batch)"
                           if dom == "gemm_tcgen05" else
                           f"conv3x3_kernel / conv3x3_rows_kernel (tcgen05 implicit GEMM, every 3x3 convolution of one {ARCH} batch)"),
                "peak_source": f"{how} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches_timed": gemm_n, "avg_launch_us": round(gemm_ms / gemm_n * 1e3, 2),
                "share_of_step_eager": breakdown.get(dom),
                "model_flops_frac_of_peak": round(GFLOP_PER_TILE * 1e9 * n_tiles_rank / (ms_per_step * 1e-3) / 1e12 / peak, 4)}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            band = synthetic_raster(ZONE_H, ZONE_W, row0=0, rows=2048)
            cpu_base, _ = cpu_baseline_sample(wpath, band, budget_s=15.0, max_tiles=24)
        except Exception as ex:  # noqa: BLE001
            cpu_base = {"value": None, "unit": "Mpx/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": f"failed: {ex!r}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mpx/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": f"{ARCH} zonal inference, synthetic {gw}x{gh}x4 uint8 @0.2m, tile {PATCH} margin {MARGIN} "
                            f"(overlap {2*MARGIN}), {len(tiles)} tiles, {N_CLS} classes, argmax raster; "
                            f"{world} row strip(s) of {ZONE_H} rows",
                "batch_tiles": args.batch, "tiles_per_gpu": n_tiles_rank, "cuda_graph": True,
                "l2": f"inputs larger than L2: {4 * in_rows * gw / 1e6:.0f} MB raster strip, >126 MB of activations per batch",
                "tiles_per_s": round(len(tiles) / (ms_per_step / 1e3), 1), "class_raster_checksum": checksum},
            "e2e": {"value": round(e2e_val, 2), "unit": "Mpx/s", "h2d_bytes_per_step": int(4 * in_rows * gw),
                    "d2h_bytes_per_step": int(res_bytes),
                    "note": "inference_and_write() on this rank's strip as a zone: pinned host raster -> HBM, fused "
                            "forward, class raster -> pinned host; file encoding excluded"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "kernel_time_shares_eager": breakdown,
        }
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

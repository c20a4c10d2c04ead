"""Per-kernel timing table of one zonal batch (eager launches, CUDA events per launch).
Writes gpurun_out/profile_batch.txt.  Diagnostic; numbers are warm-L2, back-to-back."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from flair_for_aigle_b200 import native as nv
from flair_for_aigle_b200.engine.convnext_unet import ConvNeXtCfg, ConvNeXtV2UNetEngine
from flair_for_aigle_b200.engine.zonal import ZonalRunner
from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
from flair_for_aigle_b200.synthetic import synthetic_raster, randomize_state_, DEFAULT_MEANS, DEFAULT_STDS

B = int(os.environ.get("B", "37"))
ARCH = os.environ.get("ARCH", bench.ARCH)
dev = torch.device("cuda:0")
zc = bench.zonal_config("w", "/tmp", "unused", B)
zc["monotemp_arch"] = ARCH
cfg = prepare_model_config(zc)
m = FLAIR_HUB_Model(cfg, {"AERIAL_RGBI": 512}, max_batch=B)
sd = m.state_dict(); randomize_state_(sd, 1)
m.load_state_dict(sd)
m._norm = (DEFAULT_MEANS, DEFAULT_STDS)
eng = m.to(dev).engine(bench.TASK, max_batch=B)
r = torch.from_numpy(synthetic_raster(3000, 3000)).to(dev)
plan = np.zeros((B, 6), np.int32)
for i in range(B):
    plan[i] = ((i // 4) * 384, (i % 4) * 384, (i // 4) * 384 + 64, (i % 4) * 384 + 64, 384, 384)
own = np.stack([plan[:, 2], plan[:, 2] + 384, plan[:, 3], plan[:, 3] + 384], 1).astype(np.int32)
out = torch.zeros((3000, 3000), dtype=torch.uint8, device=dev)
run = ZonalRunner(eng, 64, use_graph=False, norm=(DEFAULT_MEANS, DEFAULT_STDS))
for _ in range(2):
    run.run(r, plan, own, out)
torch.cuda.synchronize()
nv.PROFILE = []
reps = 3
for _ in range(reps):
    run.run(r, plan, own, out)
torch.cuda.synchronize()
prof, nv.PROFILE = nv.PROFILE, None
agg = collections.OrderedDict()
for fam, meta, a, b in prof:
    key = (fam, tuple(sorted(meta.items())))
    t, c = agg.get(key, (0.0, 0))
    agg[key] = (t + a.elapsed_time(b), c + 1)
PEAKS = bench.peaks()[0]
TENSOR, HBM = PEAKS.get("bf16_tflops_sustained", 1354.0) * 1e12, PEAKS.get("hbm_gbs", 6530.0) * 1e9
FMA = 148 * 128 * 2 * PEAKS.get("sm_max_mhz", 1965.0) * 1e6        # fp32 CUDA-core peak (flop/s)


def cost(fam, md):
    """-> (flop, pipe peak, algorithmic HBM bytes) of one launch; None where no model is written down."""
    if fam == "gemm_tcgen05":
        M, N, K, mode = md["M"], md["N"], md["K"], md["mode"]
        out_b = 4 if mode in (2, 3) else 2
        by = 2 * M * K + 2 * N * K * (M // 1024 if mode == 2 and "convnext" in ARCH and N * 4 == K else 1)
        by += out_b * M * N + (4 * M * N if mode == 2 else 0)
        return 2.0 * M * N * K, TENSOR, by
    if fam == "conv3x3_tcgen05":
        px = md["B"] * md["H"] * md["H"]
        out_b = 1 if md["mode"] == 2 else 2
        return 2.0 * px * 9 * md["Cin"] * md["Cout"], TENSOR, px * (2 * md["Cin"] + out_b * md["Cout"])
    if fam == "dwconv7_ln":
        n = md["B"] * md["H"] * md["H"] * md["C"]
        return 2.0 * 49 * n, FMA, 6 * n
    if fam in ("ln2d_s2d", "layernorm_rows", "merge_ln"):
        n = md.get("rows", 0) * md.get("C", 0) if "rows" in md else md.get("B", md.get("n", 0)) * md["H"] * md["H"] * md["C"]
        return 0.0, FMA, 6 * n
    if fam == "stem_ln":
        n = md["n"]
        return 2.0 * 64 * n * 128 * 128 * 128, FMA, n * (4 * 512 * 512 + 4 * 128 * 128 * 128)
    if fam == "scale_weights":
        return 0.0, FMA, 2 * md["N"] * md["K"] * (md["B"] + 1)
    if fam == "upsample2_concat":
        return 0.0, FMA, 2 * md["B"] * md["H"] * md["H"] * md["C"] * 1.6     # write + (quarter-size input, skip) read
    if fam == "cast_f32_bf16":
        return 0.0, FMA, 6 * md["n"]
    if fam == "swin_window_attn":
        H, C, n = md["H"], md["C"], md["n"]
        nw = ((H + 11) // 12) ** 2
        return n * nw * (C // 32) * 4.0 * 144 * 144 * 32, TENSOR, n * H * H * C * 2 * 4
    return None


tot = sum(t for t, _ in agg.values()) / reps
lines = [f"{ARCH} batch B={B}: total kernel time {tot:.3f} ms = {tot/B:.4f} ms/tile (eager, per-launch events)"]
fam_tot = collections.Counter()
for (fam, meta), (t, c) in agg.items():
    fam_tot[fam] += t / reps
    md = dict(meta)
    extra = ""
    if fam == "gemm_tcgen05":
        extra = f" {2.0*md['M']*md['N']*md['K']/(t/c*1e-3)/1e12:7.1f} TFLOP/s"
    if fam == "conv3x3_tcgen05":
        fl = 2.0 * md['B'] * md['H'] * md['H'] * 9 * md['Cin'] * md['Cout']
        extra = f" {fl/(t/c*1e-3)/1e12:7.1f} TFLOP/s"
    roof = ""
    cm = cost(fam, md)
    if cm is not None:
        fl, pk, by = cm
        t_pipe, t_hbm, t_meas = fl / pk, by / HBM, t / c * 1e-3
        bound = "tensor" if (pk == TENSOR and t_pipe >= t_hbm) else ("fma" if t_pipe >= t_hbm else "hbm")
        roof = f"  | roofline {max(t_pipe, t_hbm)*1e6:7.1f} us ({bound}) = {max(t_pipe, t_hbm)/t_meas*100:4.0f}%  {by/1e6:7.1f} MB"
    lines.append(f"{fam:18s} x{c//reps:3d} avg {t/c*1e3:9.1f} us  sum/batch {t/reps:8.3f} ms {extra}{roof}  {md}")
lines.append("--- by family (ms per batch, share)")
for fam, t in fam_tot.most_common():
    lines.append(f"{fam:18s} {t:8.3f} ms  {t/tot*100:5.1f}%")
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/profile_batch_{ARCH.split('_')[0].split('-')[0]}.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

"""Layer-by-layer comparison of the ConvNeXt-V2/U-Net engine with the fp32 oracle on the GPU box.
Writes gpurun_out/model_check.log.  Diagnostic only (imports oracle/, never the other way)."""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from flair_for_aigle_b200 import native as nv
from flair_for_aigle_b200.engine.convnext_unet import ConvNeXtCfg, ConvNeXtV2UNetEngine
from flair_for_aigle_b200.synthetic import synthetic_raster
from oracle.models import FlairHubOracle, randomize_

os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/model_check.log", "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    log.write(s + "\n")
    log.flush()


torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
task = "AERIAL_LABEL-COSIA"
n = int(os.environ.get("N_TILES", "2"))
impl = os.environ.get("FZ_GEMM_IMPL", "tcgen05")
try:
    oracle = FlairHubOracle("convnextv2_base-unet", {"AERIAL_RGBI": 4}, {task: 19}).eval()
    randomize_(oracle, seed=2025, bf16_exact=True)
    sd = {k: v.clone() for k, v in oracle.state_dict().items()}
    mean, std = [105.66, 111.35, 102.18, 106.59], [52.23, 45.62, 44.30, 39.78]
    eng = ConvNeXtV2UNetEngine(sd, "encoders.AERIAL_RGBI.seg_model.model.", f"main_decoders.{task}.seg_model.",
                               ConvNeXtCfg(), dev, max_batch=n, norm_mean=mean, norm_std=std)
    eng.gemm_impl = impl
    oracle = oracle.to(dev)
    raster = synthetic_raster(1100, 1100, seed=2025)
    tiles = np.stack([raster[:, (37 * i) % 500:(37 * i) % 500 + 512, (91 * i) % 500:(91 * i) % 500 + 512] for i in range(n)])
    u8 = torch.from_numpy(tiles).to(dev)
    xn = ((u8.double() - torch.tensor(mean, device=dev, dtype=torch.float64).view(1, 4, 1, 1)) /
          torch.tensor(std, device=dev, dtype=torch.float64).view(1, 4, 1, 1)).float()

    acts = {}
    enc = oracle.encoders["AERIAL_RGBI"].seg_model.model
    dec = oracle.main_decoders[task].seg_model
    hooks = [enc.stem_1.register_forward_hook(lambda m, i, o: acts.__setitem__("stem", o))]
    for si in range(4):
        hooks.append(getattr(enc, f"stages_{si}").register_forward_hook(
            lambda m, i, o, si=si: acts.__setitem__(f"stage{si}", o)))
    hooks.append(enc.stages_0.blocks[0].register_forward_hook(lambda m, i, o: acts.__setitem__("s0b0", o)))
    for k in range(5):
        hooks.append(dec.decoder.blocks[k].register_forward_hook(lambda m, i, o, k=k: acts.__setitem__(f"dec{k}", o)))
    with torch.no_grad():
        t0 = time.time()
        ref, _ = oracle({"AERIAL_RGBI": xn, task: torch.zeros(n, 19, 512, 512, device=dev)})
        torch.cuda.synchronize()
        P(f"oracle forward (torch eager fp32 on GPU): {time.time()-t0:.2f}s")
    ref = ref[task]

    def cmp(name, got_nhwc, want_nchw):
        got = got_nhwc.float().permute(0, 3, 1, 2)
        d = (got - want_nchw).abs()
        P(f"{name:8s} shape {tuple(want_nchw.shape)} ref std {want_nchw.std().item():.4f} max|d| {d.max().item():.5f} "
          f"mean|d| {d.mean().item():.6f}  rel-to-std max {d.max().item()/want_nchw.std().item():.4f}")

    # stem only
    tiles_nhwc = u8.permute(0, 2, 3, 1).contiguous()
    nv.stem_ln(tiles_nhwc, eng.stem_w_u8, eng.stem_b_u8, eng.stem_ln_w, eng.stem_ln_b, eng.x[0][:n])
    torch.cuda.synchronize()
    cmp("stem", eng.x[0][:n], acts["stem"])
    t0 = time.time()
    eng.encode_u8(tiles_nhwc)
    torch.cuda.synchronize()
    P(f"engine encode: {time.time()-t0:.3f}s")
    for si in range(4):
        cmp(f"stage{si}", eng.x[si][:n], acts[f"stage{si}"])
    # decoder, block by block (re-run body to capture intermediates)
    a = eng.x[3][:n]
    skips = [eng.x[2][:n], eng.x[1][:n], eng.x[0][:n], None, None]
    hd = eng.hw[3]
    for k, blk in enumerate(eng.dec):
        hd *= 2
        ct = blk["cin"] + blk["cskip"]
        cat = eng.decoder.cat[:n * hd * hd * ct].view(n, hd, hd, ct)
        nv.upsample2_concat(a, skips[k] if blk["cskip"] > 0 else None, cat)
        o1 = eng.decoder.t1[:n * hd * hd * blk["cout"]].view(n, hd, hd, blk["cout"])
        nv.conv3x3(cat, blk["conv1_w"], blk["conv1_s"], blk["conv1_b"], nv.CONV_RELU_BF16, out=o1)
        o2 = eng.decoder.t2[:n * hd * hd * blk["cout"]].view(n, hd, hd, blk["cout"])
        nv.conv3x3(o1, blk["conv2_w"], blk["conv2_s"], blk["conv2_b"], nv.CONV_RELU_BF16, out=o2)
        torch.cuda.synchronize()
        cmp(f"dec{k}", o2, acts[f"dec{k}"])
        a = o2
    out = eng.decode_logits_nchw(n)
    torch.cuda.synchronize()
    d = (out - ref).abs()
    sd_ = ref.std().item()
    top2 = ref.topk(2, dim=1).values
    gap = top2[:, 0] - top2[:, 1]
    agree = (out.argmax(1) == ref.argmax(1))
    P(f"logits   std {sd_:.4f} max|d| {d.max().item():.5f} mean|d| {d.mean().item():.6f} argmax agree {agree.float().mean().item():.6f}")
    for thr in (0.01, 0.02, 0.05, 0.1):
        m = gap > thr * sd_
        P(f"  pixels with oracle top-2 gap > {thr}*std: {m.float().mean().item():.4f}; agreement among them {agree[m].float().mean().item():.6f}")
    # timing of the full forward (stem..head argmax) for n tiles
    plan = torch.zeros((n, 6), dtype=torch.int32, device=dev)
    plan[:, 4:] = 384
    rast = torch.zeros((384, 384), dtype=torch.uint8, device=dev)
    for _ in range(2):
        eng.encode_u8(tiles_nhwc)
        eng.decode_argmax_to_raster(n, plan, None, rast, 64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.encode_u8(tiles_nhwc)
    eng.decode_argmax_to_raster(n, plan, None, rast, 64)
    e1.record()
    torch.cuda.synchronize()
    P(f"engine full forward n={n}: {e0.elapsed_time(e1):.3f} ms ({e0.elapsed_time(e1)/n:.3f} ms/tile) [eager launches]")
    P("MODEL_CHECK DONE")
except Exception as ex:
    P("EXC", repr(ex))
    traceback.print_exc(file=log)
    traceback.print_exc()

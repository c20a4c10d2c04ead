"""Where does the engine's logit error come from?  For seeded 512x512 tiles: the ConvNeXtV2-base + U-Net engine's four
stage outputs and logits against the fp32 oracle (measured), next to the precision simulator's prediction for the same
rounding points (tests/error_budget.py, all families in the library's operand format, engine GELU).  A stage whose
measured error is well above the simulated one has an error source the budget does not model.
Writes gpurun_out/r2_stage_errors.txt.  Test infrastructure (uses the oracle)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import error_budget as eb  # noqa: E402
from flair_for_aigle_b200 import native as nv  # noqa: E402
from flair_for_aigle_b200.engine.convnext_unet import ConvNeXtCfg, ConvNeXtV2UNetEngine  # noqa: E402

eb.no_tf32()
dev = torch.device("cuda:0")
fmt = "fp16" if nv.op_dtype() == torch.float16 else "bf16"
lines = [f"engine operand format: {fmt}"]
for seed in (1, 2, 3):
    model = eb.make_model(seed, device=dev)
    x = eb.make_tile(100 + seed, 512, dev)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    eng = ConvNeXtV2UNetEngine(sd, "encoders.AERIAL_RGBI.seg_model.model.", f"main_decoders.{eb.TASK}.seg_model.",
                               ConvNeXtCfg(), dev, max_batch=1)
    with torch.no_grad():
        ref, ref_feats = eb.simulate(model, x, {}, want_feats=True)
        sim, sim_feats = eb.simulate(model, x, {**{f: fmt for f in eb.FAMILIES}, "gelu": "fit"}, want_feats=True)
    eng.encode_f32(x)
    got_feats = [eng.x[i][:1].permute(0, 3, 1, 2).float() for i in range(4)]
    got = eng.decode_logits_nchw(1).float()
    torch.cuda.synchronize()
    lines.append(f"seed {seed}: {'tensor':8s} {'measured mean|d|/std':>22s} {'simulated':>12s} {'ratio':>7s} {'measured max|d|/std':>20s} {'simulated':>10s}")
    for name, g, s_, r in [(f"stage{i}", got_feats[i], sim_feats[i], ref_feats[i]) for i in range(4)] + [("logits", got, sim, ref)]:
        sdv = float(r.std())
        dm, ds = (g - r).abs(), (s_ - r).abs()
        lines.append(f"        {name:8s} {float(dm.mean()) / sdv:22.6f} {float(ds.mean()) / sdv:12.6f} "
                     f"{float(dm.mean()) / max(float(ds.mean()), 1e-12):7.2f} {float(dm.max()) / sdv:20.5f} {float(ds.max()) / sdv:10.5f}")
    cls = ref.argmax(1)
    lines.append(f"        class agreement with fp32: engine {float((got.argmax(1) == cls).float().mean()):.5f}, "
                 f"simulator {float((sim.argmax(1) == cls).float().mean()):.5f}; engine vs simulator "
                 f"{float((got.argmax(1) == sim.argmax(1)).float().mean()):.5f}")
    del eng, model
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/r2_stage_errors.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

"""Kernel time breakdown of one training step at configs[4] size (torch.profiler, CUDA activities)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from flair_for_aigle_b200.engine.convnext_unet import CONVNEXTV2_CFGS
from flair_for_aigle_b200.engine.train_step import ConvNeXtUNetTrainer
dev = torch.device("cuda:0")
B, P, TASK = int(os.environ.get("B", "16")), 512, "AERIAL_LABEL-COSIA"
mods = {"AERIAL_RGBI": 4, "DEM_ELEV": 1}
state = {k: v.to(dev) for k, v in bench.random_state(mods, seed=2025).items()}
depths, dims = CONVNEXTV2_CFGS["convnextv2_base"]
w = torch.ones(19, device=dev); w[15:] = 0
tr = ConvNeXtUNetTrainer(state, depths, dims, list(mods), TASK, w, cuda_graph=os.environ.get("GRAPH", "1") != "0")
g = torch.Generator(device="cpu").manual_seed(1)
batch = {k: torch.randn(B, c, P, P, generator=g).to(dev) for k, c in mods.items()}
batch[TASK] = torch.nn.functional.one_hot(torch.randint(0, 19, (B, P, P), generator=g), 19).permute(0, 3, 1, 2).float().to(dev)
tr.step(batch); tr.step(batch); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(batch); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 1e3:.1f} ms")
for e in rows[:45]:
    print(f"{e.device_time_total / 1e3:9.2f} ms {100 * e.device_time_total / tot:5.1f}%  x{e.count:5d}  {e.key[:110]}")

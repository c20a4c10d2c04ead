"""Training step at BASELINE.json configs[4] size: convnextv2_base-unet, AERIAL_RGBI (B,4,512,512) + DEM_ELEV (B,1,512,512),
weighted CE (classes 15-18 weight 0), AdamW(5e-5, wd 0.01).  Run alone or under torchrun (one rank per GPU: one NCCL
all-reduce of the flat gradient arena per step).  Correctness-first kernels: this is a first measurement, not a tuned one."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import bench
from flair_for_aigle_b200.engine.convnext_unet import CONVNEXTV2_CFGS
from flair_for_aigle_b200.engine.train_step import ConvNeXtUNetTrainer

B = int(os.environ.get("B", "16"))
P = int(os.environ.get("P", "512"))
TASK = "AERIAL_LABEL-COSIA"
mods = {"AERIAL_RGBI": 4, "DEM_ELEV": 1}
state = {k: v.to(dev) for k, v in bench.random_state(mods, seed=2025).items()}
depths, dims = CONVNEXTV2_CFGS["convnextv2_base"]
w = torch.ones(19, device=dev); w[15:] = 0
tr = ConvNeXtUNetTrainer(state, depths, dims, list(mods), TASK, w, cuda_graph=os.environ.get("GRAPH", "1") != "0")
n_par = tr.opt.arena.numel()
g = torch.Generator(device="cpu").manual_seed(2025 + rank)
batch = {k: torch.randn(B, c, P, P, generator=g).to(dev) for k, c in mods.items()}
batch[TASK] = torch.nn.functional.one_hot(torch.randint(0, 19, (B, P, P), generator=g), 19).permute(0, 3, 1, 2).float().to(dev)
losses = []
for _ in range(2):                                              # warm-up: one eager step, one that captures the graph
    loss, _ = tr.step(batch); losses.append(float(loss))
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = int(os.environ.get("STEPS", "2"))
ar = []
e0.record()
for _ in range(steps):
    loss, _ = tr.step(batch); losses.append(float(loss))
    ar.append(getattr(tr, "last_allreduce_ms", 0.0))
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
drift = torch.zeros(1, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ref = tr.opt.arena.clone()
    dist.broadcast(ref, 0)
    drift = (tr.opt.arena - ref).abs().max().reshape(1)
    dist.all_reduce(drift, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"train step convnextv2_base-unet 2 encoders, {n_par / 1e6:.1f} M parameters, per-GPU batch {B} x {P}^2, {world} GPU(s): "
          f"{ms.item():.1f} ms/step = {B * world / ms.item() * 1e3:.2f} samples/s; losses {[round(v, 4) for v in losses]}; "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB; gradient all-reduce {sum(ar) / max(len(ar), 1):.2f} ms/step; "
          f"max parameter difference across ranks {drift.item():.3g}")
if world > 1:
    dist.destroy_process_group()

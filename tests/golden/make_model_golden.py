"""Generates tests/golden/model_<arch>.npz on CPU from the REFERENCE's ``FLAIR_HUB_Model.forward`` (flair_model.py:357-430,
imported from /root/reference behind tests/reference_stubs.py; its smp / timm sub-modules, absent from the image, are the
restated ones of oracle/models.py): for one seeded 512x512 tile and seeded bf16-exact random weights in the reference's
state_dict layout, the central 256x256 of the class map, a bit mask of the pixels whose top-2 logit gap exceeds 5 % of
the logit std ("confident" pixels) and a few logit statistics.  The CPU suite checks that the oracle reproduces them, the
GPU suite checks the engines against them without the oracle at run time.
Run:  python tests/golden/make_model_golden.py          (needs /root/reference)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
HERE = os.path.dirname(os.path.abspath(__file__))
ARCHS = {"convnextv2_base-unet": 21, "swin_base_patch4_window12_384-upernet": 22, "resnet34-unet": 23}
TASK = "AERIAL_LABEL-COSIA"


def golden_inputs(arch: str, seed: int):
    """(state_dict, uint8 tile [4,512,512], normalised float tile) -- shared by the generator and the tests."""
    import bench
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS, randomize_state_, synthetic_raster
    c = bench.zonal_config("unused", "/tmp", "unused", 1)
    c["monotemp_arch"] = arch
    sd = FLAIR_HUB_Model(prepare_model_config(c), {"AERIAL_RGBI": 512}, max_batch=1).state_dict()
    randomize_state_(sd, seed=seed)
    if arch.startswith("resnet"):
        for k in sd:
            if ".conv2.weight" in k and "layer" in k:
                sd[k].mul_(0.25)
    tile = synthetic_raster(512, 512, seed=seed)
    mean = torch.tensor(DEFAULT_MEANS, dtype=torch.float64).view(4, 1, 1)
    std = torch.tensor(DEFAULT_STDS, dtype=torch.float64).view(4, 1, 1)
    xn = ((torch.from_numpy(tile).double() - mean) / std).float()[None]
    return sd, tile, xn


def zone_weights(path: str, arch: str = "resnet34-unet", seed: int = 31) -> None:
    """Checkpoint (.safetensors, reference layout) of the small-zone golden (zone_small.npz): seeded, every parameter and
    buffer randomised, activations O(1) -- shared by make_reference_golden.py and the GPU test."""
    from safetensors.torch import save_file
    sd, _, _ = golden_inputs(arch, seed)
    save_file({k: v.contiguous() for k, v in sd.items()}, path)


def oracle_logits(arch: str, sd, xn):
    from oracle.models import FlairHubOracle
    o = FlairHubOracle(arch, {"AERIAL_RGBI": 4}, {TASK: 19}).eval()
    o.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out, _ = o({"AERIAL_RGBI": xn, TASK: torch.zeros(1, 19, 512, 512)})
    return out[TASK][0]


def reference_logits(arch: str, sd, xn):
    """The reference's own model class and forward (needs /root/reference)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import reference_stubs as rs
    rs.install()
    import test_reference_pin as pin
    from flair_hub.models.flair_model import FLAIR_HUB_Model as RefModel
    m = RefModel(pin._model_cfg(arch, {"AERIAL_RGBI": 4}), {"AERIAL_RGBI": 512}).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out, _ = m({"AERIAL_RGBI": xn, TASK: torch.zeros(1, 19, 512, 512)})
    return out[TASK][0]


def summarise(logits: torch.Tensor):
    c = logits[:, 128:384, 128:384]
    top2 = c.topk(2, dim=0).values
    conf = (top2[0] - top2[1]) > 0.05 * logits.std()
    return {"classes": c.argmax(0).to(torch.uint8).numpy(), "confident": np.packbits(conf.numpy()),
            "stats": np.array([logits.mean().item(), logits.std().item(), logits.abs().max().item()], np.float64),
            "probe": logits[:, ::97, ::89].numpy().astype(np.float32)}


if __name__ == "__main__":
    torch.manual_seed(0)
    for arch, seed in ARCHS.items():
        sd, tile, xn = golden_inputs(arch, seed)
        g = summarise(reference_logits(arch, sd, xn))
        path = os.path.join(HERE, f"model_{arch.split('-')[0]}.npz")
        np.savez_compressed(path, **g)
        print(arch, "->", path, os.path.getsize(path), "bytes; stats", g["stats"], "confident frac",
              np.unpackbits(g["confident"]).mean())

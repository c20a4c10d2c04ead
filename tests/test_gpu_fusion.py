"""Two mono-temporal encoders + FusionHandler (flair_model.py:473-547) + U-Net decoder: the model of BASELINE.json
configs[4] (AERIAL_RGBI 4 ch + DEM_ELEV 1 ch), forward, against the fp32 oracle."""
import pytest
import torch
from parity import (CLASS_AGREEMENT, CLASS_AGREEMENT_CONFIDENT, CLASS_AGREEMENT_FUSED, LOGIT_MAX_ABS,  # noqa: E402,F401
                    LOGIT_MEAN_ABS)

TASK = "AERIAL_LABEL-COSIA"


def _config(dem_patch=512):
    import bench
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    c = bench.zonal_config("unused.safetensors", "/tmp", "unused", 2)
    c["modalities"]["inputs"]["DEM_ELEV"] = True
    c["modalities"]["DEM_ELEV"] = {"input_img_path": "unused", "channels": [1], "calc_elevation": True,
                                   "calc_elevation_stack_dsm": False,
                                   "normalization": {"type": "custom", "means": [0.0], "stds": [1.0]}}
    return prepare_model_config(c)


def test_state_dict_layout_matches_oracle():
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from oracle.models import FlairHubOracle
    m = FLAIR_HUB_Model(_config(), {"AERIAL_RGBI": 512, "DEM_ELEV": 512}, max_batch=2)
    o = FlairHubOracle("convnextv2_base-unet", {"AERIAL_RGBI": 4, "DEM_ELEV": 1}, {TASK: 19})
    a = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in o.state_dict().items()}
    assert a == b
    assert a["fusion_handler.conv_f.2.weight"] == (512, 1024, 1, 1)
    assert a["encoders.DEM_ELEV.seg_model.model.stem_0.weight"] == (128, 1, 4, 4)


@pytest.mark.gpu
def test_two_modality_forward_vs_oracle(cuda):
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.synthetic import randomize_state_
    from oracle.models import FlairHubOracle
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = FLAIR_HUB_Model(_config(), {"AERIAL_RGBI": 512, "DEM_ELEV": 512}, max_batch=2)
    sd = m.state_dict()
    randomize_state_(sd, seed=11)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    o = FlairHubOracle("convnextv2_base-unet", {"AERIAL_RGBI": 4, "DEM_ELEV": 1}, {TASK: 19}).eval()
    o.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    o = o.to(cuda)
    g = torch.Generator(device="cpu").manual_seed(3)
    batch = {"AERIAL_RGBI": torch.randn(2, 4, 512, 512, generator=g).to(cuda),
             "DEM_ELEV": torch.randn(2, 1, 512, 512, generator=g).to(cuda),
             TASK: torch.zeros(2, 19, 512, 512, device=cuda)}
    with torch.no_grad():
        ref, _ = o(batch)
    out, aux = m(batch)
    torch.cuda.synchronize()
    assert aux == {}
    ref, out = ref[TASK], out[TASK]
    sd_ = ref.std().item()
    d = (out - ref).abs()
    agree = (out.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"fused logits: max|d|={d.max().item():.4f} mean|d|={d.mean().item():.5f} std={sd_:.3f} agree={agree:.5f}")
    assert d.mean().item() < LOGIT_MEAN_ABS * sd_ and d.max().item() < LOGIT_MAX_ABS * sd_
    out2, _ = m(batch)
    assert torch.equal(out, out2[TASK])


@pytest.mark.gpu
def test_two_modalities_of_different_patch_size(cuda):
    """DEM at twice the aerial patch size: FusionHandler resizes its feature maps to the aerial ones (bilinear)."""
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.synthetic import randomize_state_
    from oracle.models import FlairHubOracle
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = FLAIR_HUB_Model(_config(), {"AERIAL_RGBI": 512, "DEM_ELEV": 1024}, max_batch=1)
    sd = m.state_dict()
    randomize_state_(sd, seed=12)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    o = FlairHubOracle("convnextv2_base-unet", {"AERIAL_RGBI": 4, "DEM_ELEV": 1}, {TASK: 19}).eval()
    o.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    o = o.to(cuda)
    g = torch.Generator(device="cpu").manual_seed(4)
    batch = {"AERIAL_RGBI": torch.randn(1, 4, 512, 512, generator=g).to(cuda),
             "DEM_ELEV": torch.randn(1, 1, 1024, 1024, generator=g).to(cuda),
             TASK: torch.zeros(1, 19, 512, 512, device=cuda)}
    with torch.no_grad():
        ref, _ = o(batch)
    out, _ = m(batch)
    torch.cuda.synchronize()
    ref, out = ref[TASK], out[TASK]
    sd_ = ref.std().item()
    d = (out - ref).abs()
    print(f"mixed-size fusion: max|d|={d.max().item():.4f} mean|d|={d.mean().item():.5f} std={sd_:.3f}")
    assert d.mean().item() < LOGIT_MEAN_ABS * sd_ and d.max().item() < LOGIT_MAX_ABS * sd_

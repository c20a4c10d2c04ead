"""GPU parity of the ResNet-34 + U-Net plan (BASELINE.json configs[0]: the reference's CPU-runnable case):
strided / residual conv epilogues, conv1 7x7, max-pool against torch fp32, then the engine against the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
from parity import (CLASS_AGREEMENT, CLASS_AGREEMENT_CONFIDENT, CLASS_AGREEMENT_FUSED, LOGIT_MAX_ABS,  # noqa: E402,F401
                    LOGIT_MEAN_ABS)
from flair_for_aigle_b200 import native as _nv  # noqa: E402
OP = _nv.op_dtype()      # the inference kernels' 16-bit operand format (float16; bfloat16 in the A/B build)
TASK = "AERIAL_LABEL-COSIA"


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize("H,Cin,Cout,stride", [(128, 64, 64, 1), (128, 64, 128, 2), (64, 128, 256, 2), (32, 256, 512, 2),
                                               (16, 512, 512, 1)])
def test_conv3x3_stride_and_residual(cuda, H, Cin, Cout, stride):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(H + Cin + stride)
    B = 2
    x = torch.randn(B, H, H, Cin, device=cuda).to(OP)
    w = (torch.randn(Cout, Cin, 3, 3, device=cuda) / (9 * Cin) ** 0.5).to(OP)
    scale, bias = torch.rand(Cout, device=cuda) + 0.5, torch.randn(Cout, device=cuda) * 0.2
    Ho = H // stride
    resid = torch.randn(B, Ho, Ho, Cout, device=cuda).to(OP)
    conv = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, stride=stride, padding=1)
    lin = (conv * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    wp = w.permute(0, 2, 3, 1).contiguous()
    for mode, ref in ((nv.CONV_RELU_BF16, torch.relu(lin)), (nv.CONV_BF16, lin),
                      (nv.CONV_ADD_RELU_BF16, torch.relu(lin + resid.float()))):
        out = torch.empty(B, Ho, Ho, Cout, dtype=OP, device=cuda)
        nv.conv3x3(x, wp, scale, bias, mode, out=out, stride=stride, resid=resid if mode == nv.CONV_ADD_RELU_BF16 else None)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item()
        assert err < 2 ** -8 * ref.abs().max().item() + 1e-3, (mode, err)
    # in-place residual (identity shortcut): out aliases resid
    buf = resid.clone()
    nv.conv3x3(x, wp, scale, bias, nv.CONV_ADD_RELU_BF16, out=buf, stride=stride, resid=buf)
    torch.cuda.synchronize()
    assert (buf.float() - torch.relu(lin + resid.float())).abs().max().item() < 2 ** -8 * lin.abs().max().item() + 1e-3


def test_conv7x7_and_maxpool(cuda):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(1)
    B, P = 2, 256
    x = torch.randn(B, 4, P, P, device=cuda)
    w = torch.randn(64, 4, 7, 7, device=cuda) * 0.1
    scale, bias = torch.rand(64, device=cuda) + 0.5, torch.randn(64, device=cuda) * 0.1
    ref = torch.relu(F.conv2d(x, w, None, stride=2, padding=3) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    wk = w.permute(2, 3, 1, 0).reshape(196, 64).contiguous()
    out = torch.empty(B, P // 2, P // 2, 64, dtype=OP, device=cuda)
    nv.conv7x7s2_bn_relu(x, wk, scale, bias, out)
    mp = torch.empty(B, P // 4, P // 4, 64, dtype=OP, device=cuda)
    nv.maxpool3x3s2(out, mp)
    torch.cuda.synchronize()
    assert (out.float().permute(0, 3, 1, 2) - ref).abs().max().item() < 2 ** -8 * ref.abs().max().item() + 1e-3
    mref = F.max_pool2d(out.float().permute(0, 3, 1, 2), 3, 2, 1)
    assert torch.equal(mp.float().permute(0, 3, 1, 2), mref)


def test_resnet34_unet_engine_vs_oracle(cuda):
    """Stated tolerance (bf16 activations through 16 BasicBlocks + 11 decoder convs): mean |dlogit| <= 1.5 %,
    max <= 15 % of the logit std; class agreement >= 98 % raw, >= 99.9 % where the oracle top-2 gap > 5 % std."""
    import bench
    from safetensors.torch import load_file, save_file
    from oracle.models import FlairHubOracle
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS, randomize_state_, synthetic_raster
    c = bench.zonal_config("w", "/tmp", "unused", 2)
    c["monotemp_arch"] = "resnet34-unet"
    model = FLAIR_HUB_Model(prepare_model_config(c), {"AERIAL_RGBI": 512}, max_batch=2)
    sd = model.state_dict()
    randomize_state_(sd, seed=11)
    model.load_state_dict(sd)
    oracle = FlairHubOracle("resnet34-unet", {"AERIAL_RGBI": 4}, {TASK: 19}).eval()
    oracle.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    oracle, model = oracle.to(cuda), model.to(cuda)
    raster = synthetic_raster(640, 1100, seed=3)
    u8 = torch.from_numpy(np.stack([raster[:, 0:512, 0:512], raster[:, 100:612, 500:1012]])).to(cuda)
    xn = ((u8.double() - torch.tensor(DEFAULT_MEANS, device=cuda, dtype=torch.float64).view(1, 4, 1, 1)) /
          torch.tensor(DEFAULT_STDS, device=cuda, dtype=torch.float64).view(1, 4, 1, 1)).float()
    with torch.no_grad():
        ref = oracle({"AERIAL_RGBI": xn, TASK: torch.zeros(2, 19, 512, 512, device=cuda)})[0][TASK]
        feats_ref = oracle.encoders["AERIAL_RGBI"].seg_model(xn)
    out = model({"AERIAL_RGBI": xn})[0][TASK]
    eng = model.engine(TASK)
    torch.cuda.synchronize()
    for i, (f, fr) in enumerate(zip(eng.features(2), feats_ref[1:])):
        rel = (f.float().permute(0, 3, 1, 2) - fr).abs().max().item() / fr.std().item()
        assert rel < 0.1, f"feature {i + 1} error {rel}"
    sd_ = ref.std().item()
    d = (out - ref).abs()
    same = out.argmax(1) == ref.argmax(1)
    top2 = ref.topk(2, dim=1).values
    conf = (top2[:, 0] - top2[:, 1]) > 0.05 * sd_
    print(f"resnet34-unet logits: max|d|={d.max().item():.4f} mean|d|={d.mean().item():.5f} std={sd_:.3f} "
          f"agree={same.float().mean().item():.5f} agree(confident)={same[conf].float().mean().item():.6f}")
    assert d.mean().item() < LOGIT_MEAN_ABS * sd_ and d.max().item() < LOGIT_MAX_ABS * sd_
    assert same.float().mean().item() >= 0.997 and same[conf].float().mean().item() >= CLASS_AGREEMENT_CONFIDENT


def test_resnet34_zone_through_public_api(cuda, tmp_path):
    """configs[0] through the drop-in API (build_inference_model -> inference_and_write) vs the oracle pipeline."""
    import bench
    from safetensors.torch import load_file, save_file
    from oracle.grid import Georef
    from oracle.models import FlairHubOracle
    from oracle.pipeline import run_zone
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection import inference as inf
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import build_inference_model, prepare_model_config
    from flair_for_aigle_b200.flair_zonal_detection.raster import RasterSink, ZoneRaster, register_raster
    from flair_for_aigle_b200.flair_zonal_detection.slicing import generate_patches_from_reference
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS, randomize_state_, synthetic_raster
    wpath = str(tmp_path / "resnet34_unet.safetensors")
    c = bench.zonal_config(wpath, str(tmp_path), "mem://r34", 4)
    c["monotemp_arch"] = "resnet34-unet"
    sd = FLAIR_HUB_Model(prepare_model_config(dict(c, model_weights=wpath)), {"AERIAL_RGBI": 512}).state_dict()
    randomize_state_(sd, seed=5)
    # keep activations O(1): shrink the second conv of every BasicBlock
    for k in sd:
        if ".conv2.weight" in k and "layer" in k:
            sd[k].mul_(0.25)
    save_file({k: v.contiguous() for k, v in sd.items()}, wpath)
    arr = synthetic_raster(700, 1000, seed=4)
    register_raster("mem://r34", ZoneRaster(arr, 700000.0, 6600000.0, 0.2))
    cfg = inf.initialize_geometry_and_resolutions(c)
    cfg["device"] = cuda
    model = build_inference_model(cfg, {"AERIAL_RGBI": 512}).to(cuda)
    tiles = generate_patches_from_reference(cfg, "mem://r34", None)
    ds = inf.prep_dataset(cfg, tiles, {"AERIAL_RGBI": 512})
    RasterSink.write_files = False
    outs, _ = inf.init_outputs(cfg, "mem://r34", 0)
    inf.inference_and_write(model, ds, tiles, cfg, outs, "mem://r34")
    got = outs[TASK].to_host()[0]
    oracle = FlairHubOracle("resnet34-unet", {"AERIAL_RGBI": 4}, {TASK: 19}).eval()
    oracle.load_state_dict(load_file(wpath), strict=True)
    ref, _, _ = run_zone(oracle.to(cuda), arr, Georef(700000.0, 6600000.0, 0.2, 1000, 700), 512, 64, DEFAULT_MEANS,
                         DEFAULT_STDS, TASK, 19, batch_size=2, device="cuda")
    agree = (got == ref).mean()
    print(f"resnet34-unet zone class agreement with the oracle pipeline: {agree:.5f}")
    assert agree >= CLASS_AGREEMENT

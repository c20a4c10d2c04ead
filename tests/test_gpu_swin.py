"""GPU parity of the Swin / UPerNet kernels (csrc/swin_ops.cu, csrc/upernet_ops.cu) against torch fp32 and the
oracle's window helpers (oracle/swin_upernet.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
from parity import (CLASS_AGREEMENT, CLASS_AGREEMENT_CONFIDENT, CLASS_AGREEMENT_FUSED, LOGIT_MAX_ABS,  # noqa: E402,F401
                    LOGIT_MEAN_ABS)
from flair_for_aigle_b200 import native as _nv  # noqa: E402
OP = _nv.op_dtype()      # the inference kernels' 16-bit operand format (float16; bfloat16 in the A/B build)


@pytest.mark.parametrize("rows,C", [(1000, 128), (513, 256), (64, 512), (77, 1024), (9, 2048)])
def test_layernorm_rows(cuda, rows, C):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(C)
    x = torch.randn(rows, C, device=cuda) * 3 + 1
    w, b = torch.rand(C, device=cuda) + 0.5, torch.randn(C, device=cuda) * 0.1
    out = torch.empty(rows, C, dtype=OP, device=cuda)
    nv.layernorm_rows(x, w, b, out, eps=1e-5)
    ref = F.layer_norm(x, (C,), w, b, 1e-5)
    assert (out.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,C", [(2, 16, 128), (3, 8, 256), (1, 4, 512)])
def test_merge_ln(cuda, B, H, C):
    from flair_for_aigle_b200 import native as nv
    from oracle.swin_upernet import PatchMerging
    torch.manual_seed(H)
    pm = PatchMerging(C, 2 * C).to(cuda)
    with torch.no_grad():
        pm.norm.weight.uniform_(0.5, 1.5)
        pm.norm.bias.normal_(0, 0.1)
        pm.reduction.weight.copy_(torch.eye(2 * C, 4 * C))     # look at the normalised gather itself
    x = torch.randn(B, H, H, C, device=cuda)
    out = torch.empty(B, H // 2, H // 2, 4 * C, dtype=OP, device=cuda)
    nv.merge_ln(x, pm.norm.weight, pm.norm.bias, out, eps=1e-5)
    with torch.no_grad():
        xr = x.reshape(B, H // 2, 2, H // 2, 2, C).permute(0, 1, 3, 4, 2, 5).flatten(3)
        ref = pm.norm(xr)
    assert (out.float() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


def _attn_reference(qkv, bias_bf, table, heads, ws, shift, scale):
    """timm SwinTransformerBlock._attn between the qkv and proj linears, in fp32, on bf16-rounded q/k/v.
    qkv: [B,H,W,3C] float (natural token order); padded tokens take `bias_bf` (what qkv(0) gives)."""
    from oracle.swin_upernet import relative_position_index, shift_attn_mask, window_partition, window_reverse
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    sx = torch.roll(qkv, shifts=(-shift, -shift), dims=(1, 2)) if shift else qkv
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    Hp, Wp = H + ph, W + pw
    full = bias_bf.view(1, 1, 1, C3).expand(B, Hp, Wp, C3).clone()
    full[:, :H, :W] = sx
    xw = window_partition(full, ws).view(-1, ws * ws, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = xw.unbind(0)
    attn = (q * scale) @ k.transpose(-2, -1)
    idx = relative_position_index(ws).to(qkv.device)
    bias = table.t()[idx.view(-1)].view(ws * ws, ws * ws, heads).permute(2, 0, 1)      # table is [heads][(2ws-1)^2]
    attn = attn + bias.unsqueeze(0)
    if shift:
        mask = shift_attn_mask(H, W, ws, shift).to(qkv.device)
        nw = mask.shape[0]
        attn = (attn.view(-1, nw, heads, ws * ws, ws * ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, ws * ws, ws * ws)
    out = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, ws, ws, C)
    out = window_reverse(out, ws, Hp, Wp)[:, :H, :W].contiguous()
    return torch.roll(out, shifts=(shift, shift), dims=(1, 2)) if shift else out


@pytest.mark.parametrize("B,H,W,heads,ws,shift", [
    (2, 16, 16, 2, 12, 0), (2, 16, 16, 2, 12, 6), (1, 32, 32, 4, 12, 6), (3, 24, 24, 1, 12, 6),
    (1, 128, 128, 4, 12, 6), (2, 14, 14, 3, 7, 3), (1, 20, 28, 2, 8, 4), (2, 12, 12, 8, 12, 0),
])
def test_window_attention(cuda, B, H, W, heads, ws, shift):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(H * 7 + shift)
    C = heads * 32
    qkv = (torch.randn(B, H, W, 3 * C, device=cuda) * 1.5).to(OP)
    bias_bf = (torch.randn(3 * C, device=cuda) * 0.5).to(OP)
    table = torch.randn(heads, (2 * ws - 1) ** 2, device=cuda) * 0.5
    scale = 32 ** -0.5
    out = torch.full((B, H, W, C), float("nan"), dtype=OP, device=cuda)
    nv.swin_window_attn(qkv, bias_bf, table, out, heads, ws, shift, scale)
    torch.cuda.synchronize()
    ref = _attn_reference(qkv.float(), bias_bf.float(), table, heads, ws, shift, scale)
    assert not torch.isnan(out.float()).any(), "some token was never written"
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-2 * max(1.0, ref.abs().max().item()), err      # P and the output are rounded to bf16
    out2 = torch.empty_like(out)
    nv.swin_window_attn(qkv, bias_bf, table, out2, heads, ws, shift, scale)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("S", [1, 2, 3, 6])
def test_adaptive_avgpool(cuda, S):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(S)
    x = torch.randn(3, 16, 16, 64, device=cuda).to(OP)
    out = torch.empty(3, S, S, 64, dtype=OP, device=cuda)
    nv.adaptive_avgpool(x, S, out)
    ref = F.adaptive_avg_pool2d(x.float().permute(0, 3, 1, 2), S).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() < 1e-2


@pytest.mark.parametrize("h,H", [(1, 16), (2, 16), (3, 16), (6, 16), (16, 32), (16, 128), (32, 128), (64, 128), (128, 128)])
def test_bilinear_slice(cuda, h, H):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(h * 131 + H)
    C, Ctot, c0 = 32, 80, 24
    x = torch.randn(2, h, h, C, device=cuda).to(OP)
    add = torch.randn(2, H, H, C, device=cuda).to(OP)
    out = torch.zeros(2, H, H, Ctot, dtype=OP, device=cuda)
    nv.bilinear_slice(x, out, c0, add=add)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), size=(H, H), mode="bilinear", align_corners=False)
    ref = ref.permute(0, 2, 3, 1) + add.float()
    got = out[..., c0:c0 + C].float()
    assert (got - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    assert (out[..., :c0] == 0).all() and (out[..., c0 + C:] == 0).all()


def test_updown_slice(cuda):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(5)
    x = torch.randn(2, 32, 32, 64, device=cuda).to(OP)
    out = torch.zeros(2, 32, 32, 128, dtype=OP, device=cuda)
    nv.updown_slice(x, out, 64)
    xf = x.float().permute(0, 3, 1, 2)
    up = F.interpolate(xf, size=(64, 64), mode="bilinear", align_corners=False)
    ref = F.interpolate(up, size=(32, 32), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert (out[..., 64:].float() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("H,C", [(32, 256), (24, 64), (128, 256)])
def test_pyramid_concat(cuda, H, C):
    """fz_pyramid_concat (one launch, separable resampling) against torch fp32 and against the slice-by-slice kernels
    it replaces (same values up to the fp32 operation order, i.e. at most one bf16 ulp)."""
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(H + C)
    B = 2
    p = [torch.randn(B, H >> (3 - k), H >> (3 - k), C, device=cuda).to(OP) for k in range(4)]
    got = torch.full((B, H, H, 5 * C), 7.0, dtype=OP, device=cuda)
    nv.pyramid_concat(p[0], p[1], p[2], p[3], got)
    old = torch.zeros_like(got)
    for k in range(4):
        nv.bilinear_slice(p[k], old, k * C)
    nv.updown_slice(p[3], old, 4 * C)
    assert torch.equal(got[..., 3 * C:4 * C], p[3])
    d = (got.float() - old.float()).abs()
    assert d.max().item() <= 2.0 ** -7 * max(1.0, old.float().abs().max().item())
    for k in range(3):
        ref = F.interpolate(p[k].float().permute(0, 3, 1, 2), size=(H, H), mode="bilinear", align_corners=False)
        assert (got[..., k * C:(k + 1) * C].float() - ref.permute(0, 2, 3, 1)).abs().max().item() < 2e-2 * ref.abs().max().item()
    xf = p[3].float().permute(0, 3, 1, 2)
    up = F.interpolate(xf, size=(2 * H, 2 * H), mode="bilinear", align_corners=False)
    ref = F.interpolate(up, size=(H, H), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert (got[..., 4 * C:].float() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


def test_head_upsample4(cuda):
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(6)
    lg = torch.randn(2, 32, 32, 64, device=cuda)
    out = torch.empty(2, 19, 128, 128, device=cuda)
    nv.head_upsample4(lg, 19, out)
    ref = torch.nn.UpsamplingBilinear2d(scale_factor=4)(lg[..., :19].permute(0, 3, 1, 2))
    assert (out - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("M", [37, 148, 333])
def test_gemm_fewer_rows_than_a_tile(cuda, M):
    """PSP 1x1 convs see M = batch * s^2 rows (37 at s = 1): rows beyond M are TMA zero fill, never stored."""
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(M)
    A = torch.randn(M, 1024, device=cuda).to(OP)
    W = (torch.randn(256, 1024, device=cuda) / 32).to(OP)
    bias = torch.randn(256, device=cuda)
    guard = torch.full((M + 300, 256), 7.0, dtype=OP, device=cuda)
    nv.gemm_bf16(A, W, nv.EPI_RELU_BF16, bias=bias, out=guard[:M])
    torch.cuda.synchronize()
    ref = torch.relu(A.float() @ W.float().t() + bias)
    assert (guard[:M].float() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    assert (guard[M:] == 7.0).all()


def _swin_pair(cuda, n_tiles):
    from flair_for_aigle_b200.engine.swin_upernet import SwinCfg, SwinUPerNetEngine
    from oracle.models import FlairHubOracle, randomize_
    task = "AERIAL_LABEL-COSIA"
    oracle = FlairHubOracle("swin_base_patch4_window12_384-upernet", {"AERIAL_RGBI": 4}, {task: 19}).eval()
    randomize_(oracle, seed=2025, bf16_exact=True)
    sd = {k: v.clone() for k, v in oracle.state_dict().items()}
    mean, std = [105.66, 111.35, 102.18, 106.59], [52.23, 45.62, 44.30, 39.78]
    eng = SwinUPerNetEngine(sd, "encoders.AERIAL_RGBI.seg_model.model.model.", f"main_decoders.{task}.seg_model.",
                            SwinCfg(), cuda, max_batch=n_tiles, norm_mean=mean, norm_std=std)
    return oracle.to(cuda), eng, task, mean, std


def test_swin_upernet_engine_vs_oracle(cuda):
    """End to end on 2 tiles: fp32 oracle (torch eager on the GPU, TF32 off) vs the engine.  Stated tolerance (bf16
    operands through 24 blocks + the decoder, fp32 accumulation and residual stream): stage features within 5 % of
    their std (max), logits mean |d| <= 1.5 % and max |d| <= 15 % of the logit std."""
    import numpy as np
    from flair_for_aigle_b200.synthetic import synthetic_raster
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    oracle, eng, task, mean, std = _swin_pair(cuda, 2)
    P = 512
    raster = synthetic_raster(640, 1100, seed=2025)
    u8 = torch.from_numpy(np.stack([raster[:, 0:P, 0:P], raster[:, 100:100 + P, 500:500 + P]])).to(cuda)
    xn = ((u8.double() - torch.tensor(mean, device=cuda, dtype=torch.float64).view(1, 4, 1, 1)) /
          torch.tensor(std, device=cuda, dtype=torch.float64).view(1, 4, 1, 1)).float()
    with torch.no_grad():
        ref, _ = oracle({"AERIAL_RGBI": xn, task: torch.zeros(2, 19, P, P, device=cuda)})
        ref = ref[task]
        feats_ref = oracle.encoders["AERIAL_RGBI"].seg_model(xn)[2:]
    eng.encode_u8(u8.permute(0, 2, 3, 1).contiguous())
    feats = [f.clone() for f in eng.features(2)]
    out = eng.decode_logits_nchw(2)
    torch.cuda.synchronize()
    for i, (f, fr) in enumerate(zip(feats, feats_ref)):
        rel = (f.permute(0, 3, 1, 2) - fr).abs().max().item() / fr.std().item()
        print(f"stage {i}: max err / std = {rel:.4f}")
        assert rel < 0.05, f"stage {i} feature error {rel}"
    sd_ = ref.std().item()
    d = (out - ref).abs()
    agree = (out.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"logits: max|d|={d.max().item():.4f} mean|d|={d.mean().item():.5f} std={sd_:.3f} argmax agree={agree:.5f}")
    assert d.max().item() < LOGIT_MAX_ABS * sd_ and d.mean().item() < LOGIT_MEAN_ABS * sd_
    eng.encode_f32(xn)
    out2 = eng.decode_logits_nchw(2)
    torch.cuda.synchronize()
    d2 = (out2 - ref).abs()
    assert d2.max().item() < LOGIT_MAX_ABS * sd_ and d2.mean().item() < LOGIT_MEAN_ABS * sd_
    eng.encode_f32(xn)
    out3 = eng.decode_logits_nchw(2)
    torch.cuda.synchronize()
    assert torch.equal(out2, out3)


def test_swin_upernet_zone_through_public_api(cuda, tmp_path):
    """configs[2] through the drop-in API (build_inference_model -> inference_and_write) vs the oracle pipeline."""
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
    task = bench.TASK
    wpath = str(tmp_path / "swin_upernet.safetensors")
    c = bench.zonal_config(wpath, str(tmp_path), "mem://swin", 4)
    c["monotemp_arch"] = "swin_base_patch4_window12_384-upernet"
    sd = FLAIR_HUB_Model(prepare_model_config(dict(c, model_weights=wpath)), {"AERIAL_RGBI": 512}).state_dict()
    randomize_state_(sd, seed=7)
    save_file({k: v.contiguous() for k, v in sd.items()}, wpath)
    arr = synthetic_raster(700, 1000, seed=4)
    register_raster("mem://swin", ZoneRaster(arr, 700000.0, 6600000.0, 0.2))
    cfg = inf.initialize_geometry_and_resolutions(c)
    cfg["device"] = cuda
    model = build_inference_model(cfg, {"AERIAL_RGBI": 512}).to(cuda)
    tiles = generate_patches_from_reference(cfg, "mem://swin", None)
    ds = inf.prep_dataset(cfg, tiles, {"AERIAL_RGBI": 512})
    RasterSink.write_files = False
    outs, _ = inf.init_outputs(cfg, "mem://swin", 0)
    inf.inference_and_write(model, ds, tiles, cfg, outs, "mem://swin")
    got = outs[task].to_host()[0]
    oracle = FlairHubOracle("swin_base_patch4_window12_384-upernet", {"AERIAL_RGBI": 4}, {task: 19}).eval()
    oracle.load_state_dict(load_file(wpath), strict=True)
    ref, _, _ = run_zone(oracle.to(cuda), arr, Georef(700000.0, 6600000.0, 0.2, 1000, 700), 512, 64, DEFAULT_MEANS,
                         DEFAULT_STDS, task, 19, batch_size=2, device="cuda")
    agree = (got == ref).mean()
    print(f"swin-upernet zone class agreement with the oracle pipeline: {agree:.5f}")
    assert agree >= CLASS_AGREEMENT


def test_crop_argmax_on_quarter_resolution_logits(cuda):
    """FZ_NHWC_UP4: crop/argmax with the UPerNet head's x4 bilinear evaluated in the kernel == upsample, then crop."""
    import numpy as np
    from flair_for_aigle_b200 import native as nv
    torch.manual_seed(9)
    n, P, m = 3, 128, 16
    lq = torch.randn(n, P // 4, P // 4, 64, device=cuda)
    full = torch.empty(n, 19, P, P, device=cuda)
    nv.head_upsample4(lq, 19, full)
    S = P - 2 * m
    plan = torch.tensor([[0, 0, m, m, S, S], [0, S, m, S + m, S, S - 5], [S, 0, S + m, m, S - 7, S]], dtype=torch.int32,
                        device=cuda)
    own = torch.stack([plan[:, 2], plan[:, 2] + plan[:, 4], plan[:, 3], plan[:, 3] + plan[:, 5]], 1).contiguous()
    a = torch.full((2 * P, 2 * P), 255, dtype=torch.uint8, device=cuda)
    b = torch.full((2 * P, 2 * P), 255, dtype=torch.uint8, device=cuda)
    nv.crop_argmax_write(full, nv.NCHW, m, plan, own, a)
    nv.crop_argmax_write(lq, nv.NHWC_UP4, m, plan, own, b, n_cls=19)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and (a != 255).any()

"""Pin the restated timm Swin encoder (oracle/swin_upernet.py) against HuggingFace ``SwinModel`` on the shapes
where both libraries agree by construction, and the UPerNet pieces against their definitions."""
import pytest
import torch

from oracle.swin_upernet import (PSPModule, SwinTransformer, UPerNetDecoder, relative_position_index,
                                 shift_attn_mask)


def _copy_to_hf(ours: SwinTransformer, hf):
    """our (timm-layout) weights -> HF SwinModel.  HF merges at the END of stage i, timm at the START of i+1."""
    sd = {}
    o = ours.state_dict()
    sd["embeddings.patch_embeddings.projection.weight"] = o["patch_embed.proj.weight"]
    sd["embeddings.patch_embeddings.projection.bias"] = o["patch_embed.proj.bias"]
    sd["embeddings.norm.weight"] = o["patch_embed.norm.weight"]
    sd["embeddings.norm.bias"] = o["patch_embed.norm.bias"]
    for i, stage in enumerate(ours.layers):
        if i > 0:
            sd[f"encoder.layers.{i-1}.downsample.norm.weight"] = o[f"layers.{i}.downsample.norm.weight"]
            sd[f"encoder.layers.{i-1}.downsample.norm.bias"] = o[f"layers.{i}.downsample.norm.bias"]
            sd[f"encoder.layers.{i-1}.downsample.reduction.weight"] = o[f"layers.{i}.downsample.reduction.weight"]
        for j, _ in enumerate(stage.blocks):
            p, h = f"layers.{i}.blocks.{j}.", f"encoder.layers.{i}.blocks.{j}."
            c = o[p + "norm1.weight"].numel()
            qkv_w, qkv_b = o[p + "attn.qkv.weight"], o[p + "attn.qkv.bias"]
            for k, name in enumerate(("query", "key", "value")):
                sd[h + f"attention.self.{name}.weight"] = qkv_w[k * c:(k + 1) * c]
                sd[h + f"attention.self.{name}.bias"] = qkv_b[k * c:(k + 1) * c]
            sd[h + "attention.self.relative_position_bias_table"] = o[p + "attn.relative_position_bias_table"]
            sd[h + "attention.output.dense.weight"] = o[p + "attn.proj.weight"]
            sd[h + "attention.output.dense.bias"] = o[p + "attn.proj.bias"]
            sd[h + "layernorm_before.weight"] = o[p + "norm1.weight"]
            sd[h + "layernorm_before.bias"] = o[p + "norm1.bias"]
            sd[h + "layernorm_after.weight"] = o[p + "norm2.weight"]
            sd[h + "layernorm_after.bias"] = o[p + "norm2.bias"]
            sd[h + "intermediate.dense.weight"] = o[p + "mlp.fc1.weight"]
            sd[h + "intermediate.dense.bias"] = o[p + "mlp.fc1.bias"]
            sd[h + "output.dense.weight"] = o[p + "mlp.fc2.weight"]
            sd[h + "output.dense.bias"] = o[p + "mlp.fc2.bias"]
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "relative_position_index" not in m and not m.startswith(("layernorm.", "pooler"))]
    assert not missing and not unexpected, (missing, unexpected)


@pytest.mark.parametrize("img,window,depths", [
    (256, 8, (2, 2, 2, 2)),     # no padding anywhere; shifted blocks + shift mask; last stage window = resolution
    (416, 12, (1, 1, 1, 1)),    # 104/52/26/13 padded to 108/60/36/24 in UN-shifted blocks (pad order irrelevant)
    (384, 12, (2, 2, 2, 2)),    # 96/48/24/12 divisible by 12: shifted, unpadded, window-12 rel-pos table
])
def test_swin_matches_hf(img, window, depths):
    from oracle.models import randomize_
    from transformers import SwinConfig, SwinModel
    torch.manual_seed(0)
    dim, heads = 32, (1, 2, 4, 8)
    ours = SwinTransformer(4, img, dim, depths, heads, window).eval()
    randomize_(ours, seed=11, bf16_exact=False)
    cfg = SwinConfig(image_size=img, patch_size=4, num_channels=4, embed_dim=dim, depths=list(depths),
                     num_heads=list(heads), window_size=window, mlp_ratio=4.0, qkv_bias=True, hidden_act="gelu",
                     layer_norm_eps=1e-5, drop_path_rate=0.0, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    hf = SwinModel(cfg, add_pooling_layer=False).eval()
    _copy_to_hf(ours, hf)
    x = torch.randn(2, 4, img, img)
    with torch.no_grad():
        feats = ours.forward_intermediates(x)
        out = hf(pixel_values=x, output_hidden_states=True)
    want = out.reshaped_hidden_states[-1]          # last stage output (no merge after it, no final norm)
    got = feats[-1]
    assert got.shape == want.shape
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() < 2e-4 * max(1.0, scale)


def test_relative_position_index_definition():
    ws = 12
    idx = relative_position_index(ws)
    for (i, j) in [(0, 0), (0, 143), (143, 0), (17, 95), (100, 3)]:
        yi, xi, yj, xj = i // ws, i % ws, j // ws, j % ws
        assert idx[i, j].item() == (yi - yj + ws - 1) * (2 * ws - 1) + (xi - xj + ws - 1)


def test_shift_mask_regions_on_padded_grid():
    """128 -> padded 132, window 12, shift 6: only the last window row/column mixes regions."""
    m = shift_attn_mask(128, 128, 12, 6).view(11, 11, 144, 144)
    assert (m[:10, :10] == 0).all()
    last = m[10, 3]
    # tokens of window rows 0..5 (region A) vs 6..11 (region B) must not see each other
    assert last[0, 5 * 12].item() == 0 and last[0, 6 * 12].item() == -100.0
    corner = m[10, 10]
    assert corner[0, 5].item() == 0 and corner[0, 6].item() == -100.0 and corner[0, 6 * 12].item() == -100.0


def test_upernet_shapes_and_unused_stage():
    dec = UPerNetDecoder([4, 0, 16, 32, 64, 128], pyramid_channels=24, segmentation_channels=8).eval()
    assert len(dec.fpn_stages) == 5                    # the fifth (input-resolution) block is never reached
    feats = [torch.randn(1, 4, 64, 64), torch.empty(1, 0, 32, 32), torch.randn(1, 16, 16, 16),
             torch.randn(1, 32, 8, 8), torch.randn(1, 64, 4, 4), torch.randn(1, 128, 2, 2)]
    with torch.no_grad():
        y = dec(*feats)
    assert y.shape == (1, 8, 16, 16)
    psp = PSPModule(128, 24).eval()
    with torch.no_grad():
        assert psp(feats[-1]).shape == (1, 24, 2, 2)

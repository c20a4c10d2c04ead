"""CPU: pin the oracle's restatement of the third-party model arithmetic (smp 0.4.0 / timm are not
installed) against independent implementations that ARE in this image:
  * HF transformers ConvNextV2 (embeddings, layer = block incl. GRN, stage downsample),
  * torchvision resnet34 (the base class of smp's ResNetEncoder),
  * torch.nn.functional for the U-Net decoder block.
The reference itself ships no tests or golden vectors for this path."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import models as om


def _hf_cfg(in_ch):
    from transformers import ConvNextV2Config
    return ConvNextV2Config(num_channels=in_ch, hidden_sizes=[128, 256, 512, 1024], depths=[3, 3, 27, 3])


@torch.no_grad()
def test_convnextv2_block_matches_hf():
    from transformers.models.convnextv2.modeling_convnextv2 import ConvNextV2Layer
    torch.manual_seed(0)
    dim = 128
    blk = om.ConvNeXtBlock(dim).eval()
    om.randomize_(blk, seed=3, bf16_exact=False)
    hf = ConvNextV2Layer(_hf_cfg(4), dim=dim, drop_path=0.0).eval()
    hf.dwconv.weight.copy_(blk.conv_dw.weight); hf.dwconv.bias.copy_(blk.conv_dw.bias)
    hf.layernorm.weight.copy_(blk.norm.weight); hf.layernorm.bias.copy_(blk.norm.bias)
    hf.pwconv1.weight.copy_(blk.mlp.fc1.weight); hf.pwconv1.bias.copy_(blk.mlp.fc1.bias)
    hf.grn.weight.copy_(blk.mlp.grn.weight.view(1, 1, 1, -1)); hf.grn.bias.copy_(blk.mlp.grn.bias.view(1, 1, 1, -1))
    hf.pwconv2.weight.copy_(blk.mlp.fc2.weight); hf.pwconv2.bias.copy_(blk.mlp.fc2.bias)
    x = torch.randn(2, dim, 24, 20)
    a, b = blk(x), hf(x)
    b = b[0] if isinstance(b, tuple) else b
    assert (a - b).abs().max().item() < 1e-4 * max(1.0, b.abs().max().item())


@torch.no_grad()
def test_convnextv2_stem_and_downsample_match_hf():
    from transformers.models.convnextv2.modeling_convnextv2 import ConvNextV2Embeddings, ConvNextV2Stage
    torch.manual_seed(1)
    enc = om.ConvNeXtV2Features(4, (1, 1, 1, 1), (128, 256, 512, 1024)).eval()
    om.randomize_(enc, seed=4, bf16_exact=False)
    cfg = _hf_cfg(4)
    emb = ConvNextV2Embeddings(cfg).eval()
    emb.patch_embeddings.weight.copy_(enc.stem_0.weight); emb.patch_embeddings.bias.copy_(enc.stem_0.bias)
    emb.layernorm.weight.copy_(enc.stem_1.weight); emb.layernorm.bias.copy_(enc.stem_1.bias)
    x = torch.randn(1, 4, 64, 64)
    s = enc.stem_1(enc.stem_0(x))
    assert (s - emb(x)).abs().max().item() < 1e-4
    st = ConvNextV2Stage(cfg, in_channels=128, out_channels=256, stride=2, depth=1).eval()
    ds = enc.stages_1.downsample
    hf_ds = st.downsampling_layer
    hf_ds[0].weight.copy_(ds[0].weight); hf_ds[0].bias.copy_(ds[0].bias)
    hf_ds[1].weight.copy_(ds[1].weight); hf_ds[1].bias.copy_(ds[1].bias)
    y = s
    for layer in hf_ds:
        y = layer(y)
    assert (ds(s) - y).abs().max().item() < 1e-4


def test_convnextv2_base_geometry_and_keys():
    enc = om.TimmUniversalEncoder("convnextv2_base", 4)
    assert enc.out_channels == [4, 0, 128, 256, 512, 1024]
    n = sum(p.numel() for p in enc.parameters())
    assert abs(n - 87.69e6) < 0.05e6           # HF ConvNeXtV2-base at C_in=4
    keys = list(enc.state_dict().keys())
    assert keys[0] == "model.stem_0.weight" and "model.stages_2.blocks.26.mlp.grn.weight" in keys
    assert "model.stages_1.downsample.1.weight" in keys and enc.state_dict()["model.stages_0.blocks.0.mlp.grn.weight"].shape == (512,)
    with torch.no_grad():
        feats = enc(torch.zeros(1, 4, 64, 64))
    assert [tuple(f.shape[1:]) for f in feats] == [(4, 64, 64), (0, 32, 32), (128, 16, 16), (256, 8, 8), (512, 4, 4),
                                                   (1024, 2, 2)]


@torch.no_grad()
def test_resnet34_encoder_matches_torchvision():
    import torchvision
    tv = torchvision.models.resnet34(weights=None).eval()
    enc = om.ResNetEncoder(3).eval()
    missing = enc.load_state_dict({k: v for k, v in tv.state_dict().items() if not k.startswith("fc.")}, strict=True)
    x = torch.randn(1, 3, 96, 96)
    feats = enc(x)
    y = tv.relu(tv.bn1(tv.conv1(x)))
    assert torch.equal(feats[1], y)
    y = tv.layer4(tv.layer3(tv.layer2(tv.layer1(tv.maxpool(y)))))
    assert torch.equal(feats[5], y)
    assert sum(p.numel() for p in enc.parameters()) == 21284672          # 21.28 M without fc
    assert enc.out_channels == [3, 64, 64, 128, 256, 512]


@torch.no_grad()
def test_unet_decoder_wiring():
    dec = om.UnetDecoder([4, 0, 128, 256, 512, 1024]).eval()
    om.randomize_(dec, seed=5, bf16_exact=False)
    shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    assert shapes["blocks.0.conv1.0.weight"] == (256, 1536, 3, 3)
    assert shapes["blocks.1.conv1.0.weight"] == (128, 512, 3, 3)
    assert shapes["blocks.3.conv1.0.weight"] == (32, 64, 3, 3) and shapes["blocks.4.conv2.0.weight"] == (16, 16, 3, 3)
    feats = [torch.randn(1, 4, 64, 64), torch.empty(1, 0, 32, 32), torch.randn(1, 128, 16, 16),
             torch.randn(1, 256, 8, 8), torch.randn(1, 512, 4, 4), torch.randn(1, 1024, 2, 2)]
    out = dec(*feats)
    assert out.shape == (1, 16, 64, 64)
    # block 0 by hand: nearest x2, concat skip, (conv3x3 no bias -> BN eval -> ReLU) x2
    b0 = dec.blocks[0]
    x = torch.cat([F.interpolate(feats[5], scale_factor=2, mode="nearest"), feats[4]], dim=1)
    for conv in (b0.conv1, b0.conv2):
        bn = conv[1]
        x = F.conv2d(x, conv[0].weight, None, padding=1)
        x = (x - bn.running_mean.view(1, -1, 1, 1)) / torch.sqrt(bn.running_var.view(1, -1, 1, 1) + bn.eps)
        x = torch.relu(x * bn.weight.view(1, -1, 1, 1) + bn.bias.view(1, -1, 1, 1))
    assert (b0(feats[5], feats[4]) - x).abs().max().item() < 1e-4


def test_product_spec_equals_oracle_layout():
    """State_dict contract: the product registers exactly the oracle's keys/shapes."""
    import bench
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    cfg = prepare_model_config(bench.zonal_config("w", "/tmp", "unused", 1))
    prod = FLAIR_HUB_Model(cfg, {"AERIAL_RGBI": 512}).state_dict()
    ora = om.FlairHubOracle("convnextv2_base-unet", {"AERIAL_RGBI": 4}, {bench.TASK: 19}).state_dict()
    assert list(prod.keys()) != [] and set(prod.keys()) == set(ora.keys())
    for k in ora:
        assert tuple(prod[k].shape) == tuple(ora[k].shape), k
        assert prod[k].dtype == ora[k].dtype, k

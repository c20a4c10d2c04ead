"""Per-layer-family precision budget of the ConvNeXtV2 + U-Net forward (test infrastructure; uses the oracle).

The sm_100a engine keeps fp32 accumulators, residual stream, normalisation statistics and biases; what it ROUNDS to
a 16-bit operand format are the tensors fed to the tensor cores.  This module replays the fp32 oracle with exactly
those rounding points switched on one family at a time, for bf16 (8 significand bits) and fp16 (11), and reports how
many pixels keep the class of the un-rounded fp32 forward:

  family   rounding point (engine tensor)
  -------  ---------------------------------------------------------------------------------------------
  y        dwconv7x7 + LayerNorm output = fc1's A operand                      (csrc/convnext_ops.cu dwconv_ln)
  hidden   GELU(fc1) output = fc2's A operand, also the tensor GRN's sum of squares is taken from
  w2s      per-sample GRN-scaled fc2 weights  W2 * diag(1 + gamma * Nx)        (scale_weights)
  down     LayerNorm2d output feeding the 2x2/s2 downsample convolutions      (ln2d_s2d)
  skip     stage outputs handed to the decoder as 16-bit copies
  dec      decoder convolution outputs (conv + BN + ReLU), incl. the head's input
  wup      sub-pixel merged 3x3 weights of the up-convolutions (sums of 1-4 taps, rounded once)
  gelu     (not a rounding) the fc1 epilogue's GELU: 0.5 x (1 + tanh(x Q(min(x^2, 50)))) with the fitted quadratic Q of
           csrc/ptx.cuh (|err| < 2.6e-5) and the hardware tanh.approx.f32 (about 2^-11 relative, modelled by rounding
           the tanh value to 11 significand bits) -- "fit" = the polynomial with an exact tanh, "hw" = both

``python tests/error_budget.py [--tile 512] [--seeds 3] [--device cpu|cuda]`` prints the table that DESIGN.md section 2
quotes (committed as profiles/r2_error_budget.txt); ``tests/test_error_budget.py`` asserts its conclusions at a reduced size.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FAMILIES = ("y", "hidden", "w2s", "down", "skip", "dec", "wup")
TASK = "AERIAL_LABEL-COSIA"


def rounder(fmt: str):
    if fmt == "fp32":
        return lambda t: t
    if fmt == "fp16x2":
        # a hi / lo split of the operand (two fp16 values whose sum carries ~22 significand bits): what a tensor-core GEMM
        # with its K dimension doubled would consume
        def split(t):
            hi = t.clamp(-65504.0, 65504.0).to(torch.float16).float()
            return hi + (t - hi).to(torch.float16).float()
        return split
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]
    if fmt == "fp16":
        return lambda t: t.clamp(-65504.0, 65504.0).to(dt).float()
    return lambda t: t.to(dt).float()


def gelu_engine(x, hw: bool):
    c0, c1, c2 = 7.97507880e-01, 3.70056493e-02, -3.51517274e-04
    u = torch.clamp(x * x, max=50.0)
    t = torch.tanh(x * ((c2 * u + c1) * u + c0))
    if hw:
        t = t.to(torch.float16).float()
    return 0.5 * x * (1.0 + t)


def _block(blk, x, q):
    """oracle ConvNeXtBlock.forward with the engine's rounding points; x NCHW fp32 (the residual stream)."""
    y = blk.norm(blk.conv_dw(x).permute(0, 2, 3, 1))
    y = q["y"](y)
    h = q["hidden"](q["gelu"](F.linear(y, blk.mlp.fc1.weight, blk.mlp.fc1.bias)))
    grn = blk.mlp.grn
    gx = h.norm(p=2, dim=(1, 2), keepdim=True)
    nx = gx / (gx.mean(dim=-1, keepdim=True) + grn.eps)
    s = 1.0 + grn.weight.view(1, 1, 1, -1) * nx                               # (B,1,1,4C)
    w2 = blk.mlp.fc2.weight                                                   # (C,4C)
    w2s = q["w2s"](w2.unsqueeze(0) * s.view(s.shape[0], 1, -1))               # (B,C,4C) per-sample scaled weights
    bias = F.linear(grn.bias, w2, blk.mlp.fc2.bias)                           # W2 beta + b2, fp32
    out = torch.einsum("bhwk,bck->bhwc", h, w2s) + bias
    return out.permute(0, 3, 1, 2) + x


def _encoder(enc, x, q):
    m = enc.model
    x = m.stem_1(m.stem_0(x))
    feats = []
    for i in range(m.num_stages):
        st = getattr(m, f"stages_{i}")
        if i > 0:
            ln, conv = st.downsample[0], st.downsample[1]
            x = conv(q["down"](ln(x)))
        for blk in st.blocks:
            x = _block(blk, x, q)
        feats.append(x)
    return feats


def _upconv(a, w, q):
    """conv3x3(nearest_up2(a), w, pad 1) as 4 phases of 2x2 convolutions with merged (then rounded) taps."""
    B, C, H, W = a.shape
    out = a.new_zeros((B, w.shape[0], 2 * H, 2 * W))
    rows = {0: ([w[:, :, 0], w[:, :, 1] + w[:, :, 2]], (1, 0)), 1: ([w[:, :, 0] + w[:, :, 1], w[:, :, 2]], (0, 1))}
    for py, (wr, pad_y) in rows.items():
        for px in (0, 1):
            taps = []
            for r in wr:                                   # r: (O,C,3) over kx
                if px == 0:
                    taps.append(torch.stack([r[:, :, 0], r[:, :, 1] + r[:, :, 2]], dim=-1))
                else:
                    taps.append(torch.stack([r[:, :, 0] + r[:, :, 1], r[:, :, 2]], dim=-1))
            k = q["wup"](torch.stack(taps, dim=2))          # (O,C,2,2)
            pad_x = (1, 0) if px == 0 else (0, 1)
            out[:, :, py::2, px::2] = F.conv2d(F.pad(a, (*pad_x, *pad_y)), k)
    return out


def _conv_bn_relu(seq, x, q, a_ch=None):
    conv, bn = seq[0], seq[1]
    if a_ch is None:
        y = conv(x)
    else:   # first conv of a decoder block: x = (a low-res, skip or None)
        a, skip = x
        y = _upconv(a, conv.weight[:, :a_ch], q)
        if skip is not None:
            y = y + F.conv2d(skip, conv.weight[:, a_ch:], padding=1)
    y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
    return q["dec"](F.relu(y))


def simulate(model, x, fmts: dict, want_feats: bool = False):
    """fmts: {family: 'fp32'|'bf16'|'fp16'}.  Returns logits (B,n_cls,H,W) fp32 (and the four stage outputs)."""
    q = {f: rounder(fmts.get(f, "fp32")) for f in FAMILIES}
    g = fmts.get("gelu", "exact")
    q["gelu"] = F.gelu if g == "exact" else (lambda t: gelu_engine(t, g == "hw"))
    enc = next(iter(model.encoders.values())).seg_model
    feats = _encoder(enc, x, q)
    dec = model.main_decoders[TASK].seg_model
    skips = [q["skip"](f) for f in feats]
    a = skips[3]
    sk = [skips[2], skips[1], skips[0], None, None]
    for i, blk in enumerate(dec.decoder.blocks):
        a = _conv_bn_relu(blk.conv1, (a, sk[i]), q, a_ch=a.shape[1])
        a = _conv_bn_relu(blk.conv2, a, q)
    logits = dec.segmentation_head(a)
    return (logits, feats) if want_feats else logits


def make_model(seed: int, arch: str = "convnextv2_base-unet", device="cpu", weights: str = "budget"):
    """weights='budget': oracle.models.randomize_ (this tool's own draw); 'test': the checkpoint the GPU tests and the bench
    use (bench.make_weights -> flair_for_aigle_b200.synthetic.randomize_state_), whose decoder amplifies rounding noise more."""
    from oracle.models import FlairHubOracle, randomize_
    m = FlairHubOracle(arch, {"AERIAL_RGBI": 4}, {TASK: 19})
    if weights == "test":
        import tempfile
        import bench
        from safetensors.torch import load_file
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "w.safetensors")
            bench.make_weights(path, seed=seed)
            m.load_state_dict(load_file(path), strict=True)
    else:
        randomize_(m, seed=seed)
    return m.eval().to(device)


def make_tile(seed: int, P: int, device="cpu"):
    from flair_for_aigle_b200.synthetic import DEFAULT_MEANS, DEFAULT_STDS, synthetic_raster
    t = torch.from_numpy(synthetic_raster(P, P, seed=seed)).double()
    mean = torch.tensor(DEFAULT_MEANS, dtype=torch.float64).view(4, 1, 1)
    std = torch.tensor(DEFAULT_STDS, dtype=torch.float64).view(4, 1, 1)
    return ((t - mean) / std).float()[None].to(device)


def no_tf32():
    """fp32 means fp32: cuDNN / cuBLAS would otherwise run the oracle's convolutions and matmuls in TF32 on the GPU."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@torch.no_grad()
def budget(P: int = 512, seeds=(1,), device="cpu", arch="convnextv2_base-unet"):
    """-> {row label: (class agreement, mean|d|/std, max|d|/std)} averaged over seeds."""
    rows = {}
    no_tf32()

    def add(label, vals):
        rows.setdefault(label, []).append(vals)

    for seed in seeds:
        model = make_model(seed, arch, device)
        x = make_tile(100 + seed, P, device)
        ref = simulate(model, x, {})
        want, _ = model({"AERIAL_RGBI": x, TASK: torch.zeros(1, 19, P, P, device=device)})
        assert torch.allclose(ref, want[TASK], rtol=0, atol=2e-3 * float(ref.std())), "simulator != oracle forward in fp32"
        sd = float(ref.std())
        cls = ref.argmax(1)

        def score(fm):
            out = simulate(model, x, fm)
            d = (out - ref).abs()
            return (float((out.argmax(1) == cls).float().mean()), float(d.mean()) / sd, float(d.max()) / sd)

        for fmt in ("bf16", "fp16"):
            add(f"all families {fmt}", score({f: fmt for f in FAMILIES}))
            for fam in FAMILIES:
                add(f"only {fam} {fmt}", score({fam: fmt}))
        for fam in FAMILIES:
            add(f"all fp16, {fam} bf16", score({f: ("bf16" if f == fam else "fp16") for f in FAMILIES}))
        add("decoder (skip,dec,wup) fp16, encoder bf16",
            score({f: ("fp16" if f in ("skip", "dec", "wup") else "bf16") for f in FAMILIES}))
        add("encoder (y,hidden,w2s,down) fp16, decoder bf16",
            score({f: ("bf16" if f in ("skip", "dec", "wup") else "fp16") for f in FAMILIES}))
        add("only gelu fit (exact tanh)", score({"gelu": "fit"}))
        add("only gelu fit + tanh.approx model", score({"gelu": "hw"}))
        add("all families fp16 + gelu fit + tanh.approx", score({**{f: "fp16" for f in FAMILIES}, "gelu": "hw"}))
        add("all families bf16 + gelu fit + tanh.approx", score({**{f: "bf16" for f in FAMILIES}, "gelu": "hw"}))
        top2 = ref.topk(2, dim=1).values
        gap = (top2[:, 0] - top2[:, 1]) / sd
        for thr in (0.002, 0.01, 0.05):
            add(f"(pixels with top-2 gap < {thr} std)", (float((gap < thr).float().mean()), 0.0, 0.0))
    return {k: tuple(sum(v[i] for v in vals) / len(vals) for i in range(3)) for k, vals in rows.items()}


@torch.no_grad()
def floor_table(P: int = 512, seeds=(7,), device="cpu", weights: str = "test"):
    """Where the 16-bit-operand floor lies for a given weight draw, and what lifts it: every family fp16 (the engine), then
    ONE group of families given more operand bits -- fp32, or the hi / lo fp16 split a doubled-K GEMM would carry."""
    rows = {}
    no_tf32()
    dec_fams, enc_fams = ("skip", "dec", "wup"), ("y", "hidden", "w2s", "down")
    for seed in seeds:
        model = make_model(seed, device=device, weights=weights)
        for tile in (100 + seed, 200 + seed):
            x = make_tile(tile, P, device)
            ref = simulate(model, x, {})
            sd, cls = float(ref.std()), ref.argmax(1)

            def score(fm):
                out = simulate(model, x, fm)
                d = (out - ref).abs()
                return (float((out.argmax(1) == cls).float().mean()), float(d.mean()) / sd, float(d.max()) / sd)
            base = {f: "fp16" for f in FAMILIES}
            cases = [("all families fp16 (= the engine)", base),
                     ("... decoder activations (dec) fp32", {**base, "dec": "fp32"}),
                     ("... decoder activations (dec) fp16 hi/lo split", {**base, "dec": "fp16x2"}),
                     ("... whole decoder (skip, dec, wup) fp32", {**base, **{f: "fp32" for f in dec_fams}}),
                     ("... whole decoder fp16 hi/lo split", {**base, **{f: "fp16x2" for f in dec_fams}}),
                     ("... whole encoder (y, hidden, w2s, down) fp32", {**base, **{f: "fp32" for f in enc_fams}}),
                     ("all families fp16 hi/lo split", {f: "fp16x2" for f in FAMILIES})]
            for label, fm in cases:
                rows.setdefault(label, []).append(score(fm))
            top2 = ref.topk(2, dim=1).values
            gap = (top2[:, 0] - top2[:, 1]) / sd
            rows.setdefault("(pixels with top-2 gap < 0.002 std)", []).append((float((gap < 0.002).float().mean()), 0.0, 0.0))
    return {k: tuple(sum(v[i] for v in vals) / len(vals) for i in range(3)) for k, vals in rows.items()}


def format_table(rows, P, seeds, arch):
    lines = [f"precision budget, {arch}, {len(seeds)} seeded tile(s) of {P}x{P}, random-init weights (every parameter randomised)",
             f"{'rounded tensors':48s} {'class agreement':>16s} {'mean|d|/std':>12s} {'max|d|/std':>11s}"]
    for k, (a, m, x) in rows.items():
        lines.append(f"{k:48s} {a:16.5f} {m:12.5f} {x:11.4f}")
    return "\n".join(lines)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--seeds", type=int, default=2)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--arch", default="convnextv2_base-unet")
    ap.add_argument("--floor", action="store_true", help="the 'what lifts the floor' table for both weight draws")
    a = ap.parse_args()
    seeds = tuple(range(1, a.seeds + 1))
    if a.floor:
        for weights, wseeds in (("test", (7,)), ("budget", (1,))):
            rows = floor_table(a.tile, wseeds, a.device, weights)
            print(f"what lifts the 16-bit floor, convnextv2_base-unet, weights = {weights} (seed {wseeds[0]}), 2 tiles of "
                  f"{a.tile}x{a.tile}, simulation of the engine's rounding points on the fp32 oracle")
            print(f"{'rounded tensors':52s} {'class agreement':>16s} {'mean|d|/std':>12s} {'max|d|/std':>11s}")
            for k, (ag, m, x) in rows.items():
                print(f"{k:52s} {ag:16.5f} {m:12.5f} {x:11.4f}")
            print()
    else:
        print(format_table(budget(a.tile, seeds, a.device, a.arch), a.tile, seeds, a.arch))

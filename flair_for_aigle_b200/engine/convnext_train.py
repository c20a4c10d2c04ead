"""Training-mode forward and backward of the ConvNeXt-V2 encoder and the U-Net decoder.

    block:  x -> dwconv7x7 (+b) -> LayerNorm(1e-6) -> fc1 -> GELU(erf) -> GRN -> fc2 -> + x         (timm ConvNeXtBlock, use_grn)

Parameters in the reference's layout (``conv_dw.{weight,bias}``, ``norm.{weight,bias}``, ``mlp.fc1/fc2.{weight,bias}``,
``mlp.grn.{weight,bias}``, smp's ``decoder.blocks.i.conv{1,2}.{0,1}.*``, ``segmentation_head.0.*``); activations NHWC, fp32
residual stream and accumulation; 16-bit GEMM operands: fp16 for the forward products, bf16 for the gradient products (``ACT``
below).  The Linear layers and their gradients run on the tcgen05 GEMM (``native.gemm_bf16`` / ``native.linear_backward``),
the wide decoder convolutions on csrc/conv3x3_small.cu, everything else on csrc/backward_ops.cu.  ``ConvNeXtBlockTrain``,
``ConvNeXtV2EncoderTrain``, ``Conv3x3BnReluTrain`` and ``UnetDecoderTrain`` each keep what their backward needs between the two
calls; engine/train_step.py wires them into the model-level step."""
import os
from typing import Dict

import torch

from .. import native as nv

_L = nv.lib
_P = nv._ptr
_S = nv._stream

# Format of the FORWARD activations (LayerNorm outputs, hidden tensors, decoder activations, im2col rows, stem patches) and of
# the weights they are multiplied with: IEEE fp16.  The gradients flowing backwards stay bf16 (range), and so do the weight
# copies of the data-gradient GEMMs; weight gradients multiply the two formats directly (tcgen05 kind::f16 takes A and B
# formats independently; the mma.sync tile kernel converts while staging).  Why: tests/diag/grad_precision_budget.py -- the
# bf16 rounding of the forward tensors ALONE puts the whole-model gradient cosine against fp32 autograd at 0.992 (the engine
# measured 0.985-0.990 with a bf16 forward), fp16 at 0.999.  FZ_TRAIN_ACT=bf16 restores the bf16 forward for comparison.
ACT = torch.bfloat16 if os.environ.get("FZ_TRAIN_ACT", "f16") == "bf16" else torch.float16
ACT_F16 = 1 if ACT == torch.float16 else 0


def _chk(rc: int, what: str) -> None:
    nv._check(rc, what)


class ConvNeXtBlockTrain:
    def __init__(self, params: Dict[str, torch.Tensor], eps_ln: float = 1e-6, eps_grn: float = 1e-6, params16=None):
        """params16 (optional): {"act": {...}, "bf16": {...}} -- 16-bit copies of the parameters under the same names in the
        forward format and in bf16 (the trainer casts its whole arena once per format); the two Linear weights are then used as
        they are instead of being cast here."""
        dev = params["conv_dw.weight"].device
        if dev.type != "cuda":
            raise nv.NativeError("ConvNeXtBlockTrain runs on CUDA only (no CPU fallback)")
        C = params["conv_dw.weight"].shape[0]
        self.C, self.eps_ln, self.eps_grn = C, eps_ln, eps_grn
        f32 = lambda t: t.detach().float().contiguous()
        self.w_dw = f32(params["conv_dw.weight"].reshape(C, 49).t())            # [49][C]
        self.b_dw = f32(params["conv_dw.bias"])
        self.ln_w, self.ln_b = f32(params["norm.weight"]), f32(params["norm.bias"])
        p16 = params16 or {}
        pa, pb = p16.get("act", {}), p16.get("bf16", {})
        act = lambda k: pa[k] if k in pa else params[k].detach().to(ACT).contiguous()
        b16 = lambda k: pb[k] if k in pb else params[k].detach().to(torch.bfloat16).contiguous()
        self.w1, self.w1_b = act("mlp.fc1.weight"), b16("mlp.fc1.weight")                # [4C][C]: forward copy, backward copy
        self.b1 = f32(params["mlp.fc1.bias"])
        self.grn_w, self.grn_b = f32(params["mlp.grn.weight"]), f32(params["mlp.grn.bias"])
        self.w2, self.w2_b = act("mlp.fc2.weight"), b16("mlp.fc2.weight")                # [C][4C]
        self.b2 = f32(params["mlp.fc2.bias"])
        self.saved = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x fp32 [B,H,W,C] -> y fp32 [B,H,W,C]; keeps what backward() needs."""
        B, H, W, C = x.shape
        M, C4, dev = B * H * W, 4 * C, x.device
        x = x.contiguous()
        u = torch.empty_like(x)
        _chk(_L().fz_dwconv7_f32(_P(x), _P(self.w_dw), _P(self.b_dw), _P(u), B, H, W, C, 0, _S()), "fz_dwconv7_f32")
        # forward operands in ACT (used once, by the GEMM right after); what the backward keeps are bf16 copies written by the
        # same kernels: the weight-gradient GEMMs multiply them with bf16 gradients and one MMA takes one operand format
        a1 = torch.empty((M, C), dtype=ACT, device=dev)
        a1b = torch.empty((M, C), dtype=torch.bfloat16, device=dev) if ACT_F16 else None
        mean = torch.empty(M, dtype=torch.float32, device=dev)
        rstd = torch.empty(M, dtype=torch.float32, device=dev)
        _chk(_L().fz_layernorm_fwd_stats(_P(u), _P(self.ln_w), _P(self.ln_b), _P(a1), _P(a1b), _P(mean), _P(rstd), M, C,
                                         self.eps_ln, ACT_F16, _S()), "fz_layernorm_fwd_stats")
        h = nv.gemm_bf16(a1, self.w1, nv.EPI_BF16, bias=self.b1)                          # pre-GELU, 16-bit [M,4C]
        g, dg = torch.empty_like(h), torch.empty_like(h)                  # GELU(h) and GELU'(h); h itself is not kept
        sumsq = torch.empty((B, C4), dtype=torch.float32, device=dev)
        _chk(_L().fz_gelu_fwd_sumsq(_P(h), _P(g), _P(dg), _P(sumsq), B, H * W, C4, ACT_F16, _S()), "fz_gelu_fwd_sumsq")   # + GRN sums
        gx, nx = torch.empty_like(sumsq), torch.empty_like(sumsq)
        mu = torch.empty(B, dtype=torch.float32, device=dev)
        a2 = torch.empty_like(g)
        a2b = torch.empty(g.shape, dtype=torch.bfloat16, device=dev) if ACT_F16 else None
        _chk(_L().fz_grn_train_forward(_P(g), _P(sumsq), _P(self.grn_w), _P(self.grn_b), _P(gx), _P(nx), _P(mu), _P(a2),
                                       _P(a2b), B, H * W, C4, self.eps_grn, ACT_F16, _S()), "fz_grn_train_forward")
        y = nv.gemm_bf16(a2, self.w2, nv.EPI_RESID_F32, bias=self.b2, resid=x.view(M, C))
        self.saved = (x, u, a1b if ACT_F16 else a1, mean, rstd, dg, g, gx, nx, mu, a2b if ACT_F16 else a2)
        return y.view(B, H, W, C)

    def backward(self, dy: torch.Tensor):
        """dy fp32 [B,H,W,C] -> (dx fp32 [B,H,W,C], {parameter name: gradient in the parameter's own shape})."""
        if self.saved is None:
            raise RuntimeError("backward() before forward()")
        x, u, a1, mean, rstd, dg, g, gx, nx, mu, a2 = self.saved
        B, H, W, C = x.shape
        M, C4, dev = B * H * W, 4 * C, x.device
        dy = dy.contiguous()
        dyb = torch.empty((M, C), dtype=torch.bfloat16, device=dev)
        nv.cast_f32_bf16(dy.view(M, C), dyb)
        da2, dw2, db2 = nv.linear_backward(dyb, a2, self.w2_b)                            # fc2
        s1 = torch.empty((B, C4), dtype=torch.float32, device=dev)
        s0 = torch.empty_like(s1)
        _chk(_L().fz_sample_colreduce2(_P(da2), _P(g), _P(s1), _P(s0), B, H * W, C4, ACT_F16, _S()), "fz_sample_colreduce2")
        ca, cb = torch.empty_like(s1), torch.empty_like(s1)
        dgrn_w = torch.empty(C4, dtype=torch.float32, device=dev)
        dgrn_b = torch.empty_like(dgrn_w)
        dh = torch.empty(g.shape, dtype=torch.bfloat16, device=dev)
        db1 = torch.empty(C4, dtype=torch.float32, device=dev)
        _chk(_L().fz_grn_gelu_backward_saved(_P(da2), _P(g), _P(dg), _P(s1), _P(s0), _P(gx), _P(nx), _P(mu), _P(self.grn_w),
                                             _P(ca), _P(cb), _P(dgrn_w), _P(dgrn_b), _P(dh), _P(db1), B, H * W, C4, self.eps_grn,
                                             ACT_F16, _S()), "fz_grn_gelu_backward_saved")
        da1, dw1, db1 = nv.linear_backward(dh, a1, self.w1_b, db=db1)                       # fc1 (bias gradient from the kernel above)
        blocks = max(1, min(592, (M + 7) // 8))
        du = torch.empty_like(u)
        partial = torch.empty((blocks, 2, C), dtype=torch.float32, device=dev)
        dln = torch.empty((2, C), dtype=torch.float32, device=dev)
        _chk(_L().fz_layernorm_bwd(_P(da1), _P(u), _P(mean), _P(rstd), _P(self.ln_w), _P(du), _P(partial), _P(dln), M, C,
                                   blocks, _S()), "fz_layernorm_bwd")
        dx = torch.empty_like(x)                                                          # dx = dy + dwconv^T(du), one pass
        # (a bf16 copy of dx written by the same kernel was tried: the extra stores cost the issue-bound kernel more than the
        # cast pass of the next block saves)
        _chk(_L().fz_dwconv7_f32_add(_P(du), _P(self.w_dw), None, _P(dy), _P(dx), B, H, W, C, 1, _S()), "fz_dwconv7_f32_add")
        dw_dw = torch.empty((49, C), dtype=torch.float32, device=dev)
        db_dw = torch.empty(C, dtype=torch.float32, device=dev)
        _chk(_L().fz_dwconv7_wgrad(_P(x), _P(du), _P(dw_dw), _P(db_dw), B, H, W, C, _S()), "fz_dwconv7_wgrad")
        grads = {"conv_dw.weight": dw_dw.t().reshape(C, 1, 7, 7), "conv_dw.bias": db_dw,
                 "norm.weight": dln[0], "norm.bias": dln[1],
                 "mlp.fc1.weight": dw1, "mlp.fc1.bias": db1, "mlp.grn.weight": dgrn_w, "mlp.grn.bias": dgrn_b,
                 "mlp.fc2.weight": dw2, "mlp.fc2.bias": db2}
        return dx, grads


class _LayerNormTrain:
    """LayerNorm over C of [M,C] fp32 rows with saved statistics; bf16 and/or fp32 output."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6):
        self.w, self.b, self.eps = weight.detach().float().contiguous(), bias.detach().float().contiguous(), eps
        self.saved = None

    def forward(self, x: torch.Tensor, want_bf16: bool, want_f32: bool):
        M, C = x.shape
        dev = x.device
        ob = torch.empty((M, C), dtype=ACT, device=dev) if want_bf16 else None             # "bf16" = the 16-bit forward format
        of = torch.empty((M, C), dtype=torch.float32, device=dev) if want_f32 else None
        mean = torch.empty(M, dtype=torch.float32, device=dev)
        rstd = torch.empty(M, dtype=torch.float32, device=dev)
        _chk(_L().fz_layernorm_fwd_stats2(_P(x), _P(self.w), _P(self.b), _P(ob), _P(of), _P(mean), _P(rstd), M, C, self.eps,
                                          ACT_F16, _S()), "fz_layernorm_fwd_stats2")
        self.saved = (x, mean, rstd)
        return ob, of

    def backward(self, dy_bf16: torch.Tensor):
        """dy bf16 [M,C] -> (dx fp32 [M,C], dweight [C], dbias [C])."""
        x, mean, rstd = self.saved
        M, C = x.shape
        blocks = max(1, min(592, (M + 7) // 8))
        dx = torch.empty_like(x)
        partial = torch.empty((blocks, 2, C), dtype=torch.float32, device=x.device)
        d = torch.empty((2, C), dtype=torch.float32, device=x.device)
        _chk(_L().fz_layernorm_bwd(_P(dy_bf16), _P(x), _P(mean), _P(rstd), _P(self.w), _P(dx), _P(partial), _P(d), M, C,
                                   blocks, _S()), "fz_layernorm_bwd")
        return dx, d[0], d[1]


def _to_bf16(x_f32: torch.Tensor) -> torch.Tensor:
    out = torch.empty(x_f32.shape, dtype=torch.bfloat16, device=x_f32.device)
    nv.cast_f32_bf16(x_f32, out)
    return out


def _to_act(x_f32: torch.Tensor) -> torch.Tensor:
    """fp32 -> the forward activation format (ACT)."""
    out = torch.empty(x_f32.shape, dtype=ACT, device=x_f32.device)
    nv.cast_f32_bf16(x_f32, out)
    return out


class ConvNeXtV2EncoderTrain:
    """Training forward + backward of the whole ConvNeXt-V2 feature extractor (timm ``convnextv2_*`` under smp's
    TimmUniversalEncoder: ``stem_0`` conv4x4/s4, ``stem_1`` LayerNorm2d, ``stages_i.downsample.{0,1}`` LayerNorm2d + conv2x2/s2,
    ``stages_i.blocks.j``): fp32 NCHW normalised tiles in, the four stage outputs (fp32 NHWC) out; ``backward`` takes the
    gradients at the four outputs and returns every parameter's gradient under the reference's state_dict keys."""

    def __init__(self, params: Dict[str, torch.Tensor], depths, dims, params16=None):
        self.depths, self.dims = tuple(depths), tuple(dims)
        params16 = params16 or {}
        w = params["stem_0.weight"]
        self.cin = w.shape[1]
        self.kpad = ((self.cin * 16 + 63) // 64) * 64
        dev = w.device
        w2 = torch.zeros((dims[0], self.kpad), dtype=ACT, device=dev)
        w2[:, :self.cin * 16] = w.detach().reshape(dims[0], -1).to(ACT)
        self.stem_w, self.stem_b = w2.contiguous(), params["stem_0.bias"].detach().float().contiguous()
        self.stem_ln = _LayerNormTrain(params["stem_1.weight"], params["stem_1.bias"])
        self.down, self.blocks = [], []
        for i, (d, c) in enumerate(zip(depths, dims)):
            if i > 0:
                p = f"stages_{i}.downsample."
                wd = params[p + "1.weight"].detach()                                  # [c, c_prev, 2, 2]
                wd2 = wd.permute(0, 2, 3, 1).reshape(c, -1)
                self.down.append((_LayerNormTrain(params[p + "0.weight"], params[p + "0.bias"]),
                                  (wd2.to(ACT).contiguous(), wd2.to(torch.bfloat16).contiguous()),    # forward / backward copies
                                  params[p + "1.bias"].detach().float().contiguous()))
            else:
                self.down.append(None)
            sub = lambda src, pre: {k[len(pre):]: v for k, v in src.items() if k.startswith(pre)}
            sub16 = lambda pre: {fmt: sub(d16, pre) for fmt, d16 in params16.items()}
            self.blocks.append([ConvNeXtBlockTrain(sub(params, f"stages_{i}.blocks.{j}."),
                                                   params16=sub16(f"stages_{i}.blocks.{j}.")) for j in range(d)])
        self.saved = None

    def forward(self, x_nchw: torch.Tensor):
        B, Cin, P, _ = x_nchw.shape
        dev = x_nchw.device
        q = P // 4
        patches = torch.empty((B * q * q, self.kpad), dtype=ACT, device=dev)
        _chk(_L().fz_patchify4_nchw(_P(x_nchw.float().contiguous()), _P(patches), B, Cin, P, self.kpad, ACT_F16, _S()),
             "fz_patchify4_nchw")
        u0 = nv.gemm_bf16(patches, self.stem_w, nv.EPI_F32, bias=self.stem_b)             # conv4x4/s4 as a GEMM, fp32
        _, x = self.stem_ln.forward(u0, want_bf16=False, want_f32=True)
        h = q
        feats, geo = [], []
        for i, c in enumerate(self.dims):
            if i > 0:
                ln, wd, bd = self.down[i]
                cp = self.dims[i - 1]
                a, _ = ln.forward(x.view(-1, cp), want_bf16=True, want_f32=False)
                a4 = torch.empty((B * (h // 2) * (h // 2), 4 * cp), dtype=ACT, device=dev)
                _chk(_L().fz_s2d_bf16(_P(a), _P(a4), B, h, h, cp, 2, 0, _S()), "fz_s2d_bf16")
                geo.append((h, cp, a4))
                h //= 2
                x = nv.gemm_bf16(a4, wd[0], nv.EPI_F32, bias=bd)
            else:
                geo.append(None)
            x = x.view(B, h, h, c)
            for blk in self.blocks[i]:
                x = blk.forward(x)
            feats.append(x)
        self.saved = (B, patches, geo)
        return feats

    def backward(self, dfeats, emit=None):
        """-> {parameter name: gradient}.  ``emit(group)`` (optional) is called with the gradients of each finished stage
        (deepest first), then with the stem's: the trainer copies them into the gradient arena and starts that bucket's
        all-reduce while the shallower stages are still running."""
        B, patches, geo = self.saved
        grads: Dict[str, torch.Tensor] = {}
        d = None
        for i in reversed(range(len(self.dims))):
            c = self.dims[i]
            df = dfeats[i].contiguous()
            if d is None:
                d = df
            else:
                s = torch.empty_like(df)
                _chk(_L().fz_add_f32(_P(d.contiguous()), _P(df), _P(s), s.numel(), _S()), "fz_add_f32")
                d = s
            for j in reversed(range(self.depths[i])):
                d, g = self.blocks[i][j].backward(d)
                for k, v in g.items():
                    grads[f"stages_{i}.blocks.{j}.{k}"] = v
            if i > 0:
                ln, wd, bd = self.down[i]
                h, cp, a4 = geo[i]
                da4, dw, db = nv.linear_backward(_to_bf16(d.view(-1, c)), a4, wd[1])
                grads[f"stages_{i}.downsample.1.weight"] = dw.view(c, 2, 2, cp).permute(0, 3, 1, 2).contiguous()
                grads[f"stages_{i}.downsample.1.bias"] = db
                da = torch.empty((B * h * h, cp), dtype=torch.bfloat16, device=da4.device)
                _chk(_L().fz_s2d_bf16(_P(da4), _P(da), B, h, h, cp, 2, 1, _S()), "fz_s2d_bf16")
                dx, dg, dbeta = ln.backward(da)
                grads[f"stages_{i}.downsample.0.weight"], grads[f"stages_{i}.downsample.0.bias"] = dg, dbeta
                d = dx.view(B, h, h, cp)
            if emit is not None:
                emit({k: v for k, v in grads.items() if k.startswith(f"stages_{i}.")})
        du0, dg, dbeta = self.stem_ln.backward(_to_bf16(d.view(-1, self.dims[0])))
        grads["stem_1.weight"], grads["stem_1.bias"] = dg, dbeta
        du0b = _to_bf16(du0)
        dw = nv.weight_gradient(du0b, patches)                                                # [C0, Kpad], operands read in place
        grads["stem_0.weight"] = dw[:, :self.cin * 16].reshape(self.dims[0], self.cin, 4, 4).contiguous()
        grads["stem_0.bias"] = nv.colsum_bf16(du0b)
        if emit is not None:
            emit({k: v for k, v in grads.items() if k.startswith("stem_")})
        return grads


def _ceil64(n: int) -> int:
    return (n + 63) // 64 * 64


def _ceil16(n: int) -> int:
    return (n + 15) // 16 * 16


class Conv3x3BnReluTrain:
    """smp ``Conv2dReLU``: conv3x3 (no bias) -> BatchNorm2d with batch statistics -> ReLU, NHWC bf16 in / out.  With
    ``bn=False`` it is the segmentation head: conv3x3 + bias, fp32 output, no normalisation.

    Two convolution paths.  The wide, few-channel layers (16-64 channels in, <= 32 out, H % 8 == 0, W % 32 == 0: the 256^2 and
    512^2 levels, where the layer is HBM-bound) run csrc/conv3x3_small.cu: forward, data gradient (the same kernel on the
    flipped, transposed weights) and weight gradient from a halo'd tile staged once in shared memory -- no im2col matrix, the
    output only as wide as the layer (``npad`` = 16 or 32).  Every other shape is a tcgen05 GEMM over an explicit im2col
    (K and N zero padded to multiples of 64)."""

    BN_EPS = 1e-5

    def __init__(self, weight: torch.Tensor, bn_weight=None, bn_bias=None, bias=None, running=None):
        """running: optional (running_mean, running_var, num_batches_tracked) buffers, updated in place by forward()."""
        self.running = running
        self.cout, self.cin = weight.shape[0], weight.shape[1]
        self.kpad, self.npad = _ceil64(9 * self.cin), _ceil64(self.cout)
        dev = weight.device
        w2 = torch.zeros((self.npad, self.kpad), dtype=torch.bfloat16, device=dev)
        w2[:self.cout, :9 * self.cin] = weight.detach().permute(0, 2, 3, 1).reshape(self.cout, -1).to(torch.bfloat16)
        self.w2 = w2.contiguous()                                                 # bf16: the data-gradient GEMM's operand
        w2a = torch.zeros((self.npad, self.kpad), dtype=ACT, device=dev)
        w2a[:self.cout, :9 * self.cin] = weight.detach().permute(0, 2, 3, 1).reshape(self.cout, -1).to(ACT)
        self.w2_act = w2a.contiguous()                                            # forward format: the forward GEMM's operand
        self.bn = bn_weight is not None
        if self.bn:
            self.g, self.b = bn_weight.detach().float().contiguous(), bn_bias.detach().float().contiguous()
        else:
            self.bias = torch.zeros(self.npad, dtype=torch.float32, device=dev)
            self.bias[:self.cout] = bias.detach().float()
        # the tile kernels' weights: [tap][co][ci] for the forward, [8 - tap][ci][co] for the data gradient
        self.coutp = _ceil16(self.cout)
        self.small_ok = self.cin in (16, 32, 48, 64) and self.coutp in (16, 32)
        if self.small_ok:
            wb = weight.detach().to(torch.bfloat16)
            wf = torch.zeros((9, self.coutp, self.cin), dtype=ACT, device=dev)
            wf[:, :self.cout] = weight.detach().to(ACT).permute(2, 3, 0, 1).reshape(9, self.cout, self.cin)
            wd = torch.zeros((9, self.cin, self.coutp), dtype=torch.bfloat16, device=dev)
            wd[:, :, :self.cout] = wb.flip(2, 3).permute(2, 3, 1, 0).reshape(9, self.cin, self.cout)
            self.w_fwd, self.w_dgrad = wf.contiguous(), wd.contiguous()
        self.saved = None

    def _small(self, H: int, W: int) -> bool:
        return self.small_ok and bool(_L().fz_conv3x3_small_supported(H, W, self.cin, self.coutp))

    def forward(self, x: torch.Tensor):
        B, H, W, C = x.shape
        M, dev = B * H * W, x.device
        small = self._small(H, W)
        self.npad = self.coutp if small else _ceil64(self.cout)               # width of the conv output / its gradient
        x = x.contiguous()
        if small:
            col = None
            conv = torch.empty((M, self.npad), dtype=torch.float32, device=dev)
            _chk(_L().fz_conv3x3_small_forward(_P(x), _P(self.w_fwd), None if self.bn else _P(self.bias), _P(conv), 0, B, H, W,
                                               C, self.npad, self.npad, self.npad, ACT_F16, _S()), "fz_conv3x3_small_forward")
        else:
            col = torch.empty((M, self.kpad), dtype=ACT, device=dev)
            _chk(_L().fz_im2col3x3_bf16(_P(x), _P(col), B, H, W, C, self.kpad, _S()), "fz_im2col3x3_bf16")   # a 16-bit gather
            conv = nv.gemm_bf16(col, self.w2_act, nv.EPI_F32, bias=None if self.bn else self.bias)  # fp32 [M, npad]
        keep = x if small else col
        if not self.bn:
            self.saved = (keep, small, (B, H, W))
            return conv.view(B, H, W, self.npad)
        chunks = max(1, min(1184, M // 64))
        y = torch.empty((M, self.cout), dtype=ACT, device=dev)
        mean = torch.empty(self.cout, dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        ws = torch.empty((chunks + 1) * 2 * self.cout, dtype=torch.float32, device=dev)
        _chk(_L().fz_bn_relu_train_forward(_P(conv), self.npad, _P(self.g), _P(self.b), _P(y), _P(mean), _P(rstd), _P(ws), M,
                                           self.cout, chunks, self.BN_EPS, ACT_F16, _S()), "fz_bn_relu_train_forward")
        if self.running is not None:
            rm, rv, nbt = self.running
            _chk(_L().fz_bn_update_running(_P(mean), _P(rstd), _P(rm), _P(rv), self.cout, M, self.BN_EPS, 0.1, _S()),
                 "fz_bn_update_running")
            if nbt is not None:
                nbt.add_(1)
        self.saved = (keep, small, conv, y, mean, rstd, chunks, (B, H, W))
        return y.view(B, H, W, self.cout)

    def backward(self, dy: torch.Tensor):
        """BN variant: dy bf16 [B,H,W,cout]; head variant: dy bf16 [B,H,W,npad] (padding columns zero).
        -> (dx fp32 [B,H,W,cin], {weight, bn_weight, bn_bias | bias})."""
        grads = {}
        if self.bn:
            keep, small, conv, y, mean, rstd, chunks, (B, H, W) = self.saved
            M, dev = B * H * W, keep.device
            dconv = torch.empty((M, self.npad), dtype=torch.bfloat16, device=dev)
            dgb = torch.empty((2, self.cout), dtype=torch.float32, device=dev)
            ws = torch.empty((chunks + 1) * 2 * self.cout, dtype=torch.float32, device=dev)
            _chk(_L().fz_bn_relu_backward(_P(conv), self.npad, _P(dy.contiguous()), _P(y), _P(mean), _P(rstd), _P(self.g),
                                          _P(dconv), self.npad, _P(dgb), _P(ws), M, self.cout, chunks, _S()),
                 "fz_bn_relu_backward")
            grads["bn_bias"], grads["bn_weight"] = dgb[0], dgb[1]
        else:
            keep, small, (B, H, W) = self.saved
            dconv = dy.contiguous().view(-1, self.npad)
        dx = torch.empty((B, H, W, self.cin), dtype=torch.float32, device=dconv.device)
        if small:
            dw = torch.empty((9, self.npad, self.cin), dtype=torch.float32, device=dconv.device)
            _chk(_L().fz_conv3x3_small_wgrad(_P(keep), _P(dconv), self.npad, _P(dw), B, H, W, self.cin, self.npad, ACT_F16,
                                             _S()), "fz_conv3x3_small_wgrad")
            grads["weight"] = dw[:, :self.cout].permute(1, 2, 0).reshape(self.cout, self.cin, 3, 3).contiguous()
            if not self.bn:
                grads["bias"] = nv.colsum_bf16(dconv)[:self.cout]
            _chk(_L().fz_conv3x3_small_forward(_P(dconv), _P(self.w_dgrad), None, _P(dx), 0, B, H, W, self.npad, self.cin,
                                               self.cin, self.cin, 0, _S()), "fz_conv3x3_small_forward")
            return dx, grads
        dcol, dw, db = nv.linear_backward(dconv, keep, self.w2)
        if not self.bn:
            grads["bias"] = db[:self.cout]
        grads["weight"] = dw[:self.cout, :9 * self.cin].reshape(self.cout, 3, 3, self.cin).permute(0, 3, 1, 2).contiguous()
        _chk(_L().fz_col2im3x3(_P(dcol), _P(dx), B, H, W, self.cin, self.kpad, _S()), "fz_col2im3x3")
        return dx, grads


class UnetDecoderTrain:
    """smp 0.4.0 ``UnetDecoder`` + ``SegmentationHead`` in training mode for the transformer-style feature list
    [x, 0-channel dummy, f4, f8, f16, f32]: five blocks (nearest x2, concat the skip when it exists, two conv-BN-ReLU), then
    conv3x3 + bias to the class logits.  Parameters under ``decoder.blocks.i.conv{1,2}.{0,1}.*`` / ``segmentation_head.0.*``."""

    def __init__(self, params: Dict[str, torch.Tensor], n_blocks: int = 5):
        self.blocks = []
        for i in range(n_blocks):
            p = f"decoder.blocks.{i}."
            def running(k):
                q = p + f"conv{k}.1."
                if q + "running_mean" not in params:
                    return None
                return (params[q + "running_mean"], params[q + "running_var"], params.get(q + "num_batches_tracked"))
            self.blocks.append(tuple(Conv3x3BnReluTrain(params[p + f"conv{k}.0.weight"], params[p + f"conv{k}.1.weight"],
                                                        params[p + f"conv{k}.1.bias"], running=running(k)) for k in (1, 2)))
        self.head = Conv3x3BnReluTrain(params["segmentation_head.0.weight"], bias=params["segmentation_head.0.bias"])
        self.n_classes = params["segmentation_head.0.weight"].shape[0]
        self.saved = None

    def forward(self, feats):
        """feats: [f4, f8, f16, f32] NHWC (fp32 or bf16) -> logits fp32 [B, n_classes, H, W]."""
        skips = list(feats[:-1][::-1]) + [None] * (len(self.blocks) - len(feats) + 1)
        x = feats[-1]
        shapes = []
        for (c1, c2), skip in zip(self.blocks, skips):
            B, H, W, C1 = x.shape
            C2 = 0 if skip is None else skip.shape[-1]
            cat = torch.empty((B, 2 * H, 2 * W, C1 + C2), dtype=ACT, device=x.device)
            nv.upsample2_concat(x.contiguous(), None if skip is None else skip.contiguous(), cat)
            shapes.append((B, H, W, C1, C2))
            x = c2.forward(c1.forward(cat))
        lg = self.head.forward(x)                                               # [B,H,W,npad] fp32
        self.saved = shapes
        return lg[..., :self.n_classes].permute(0, 3, 1, 2).contiguous()

    def backward(self, dlogits: torch.Tensor):
        """dlogits fp32 [B, n_classes, H, W] -> ([df4, df8, df16, df32] fp32 NHWC, parameter gradients)."""
        shapes = self.saved
        B, C, H, W = dlogits.shape
        dl = torch.zeros((B, H, W, self.head.npad), dtype=torch.bfloat16, device=dlogits.device)
        dl[..., :C] = dlogits.permute(0, 2, 3, 1).to(torch.bfloat16)
        grads: Dict[str, torch.Tensor] = {}
        d, g = self.head.backward(dl)
        grads["segmentation_head.0.weight"], grads["segmentation_head.0.bias"] = g["weight"], g["bias"]
        dskips = []
        for i in reversed(range(len(self.blocks))):
            c1, c2 = self.blocks[i]
            for conv, k in ((c2, 2), (c1, 1)):
                d, g = conv.backward(_to_bf16(d))
                p = f"decoder.blocks.{i}.conv{k}."
                grads[p + "0.weight"], grads[p + "1.weight"], grads[p + "1.bias"] = g["weight"], g["bn_weight"], g["bn_bias"]
            Bq, Hq, Wq, C1, C2 = shapes[i]
            da = torch.empty((Bq, Hq, Wq, C1), dtype=torch.float32, device=d.device)
            ds = torch.empty((Bq, 2 * Hq, 2 * Wq, C2), dtype=torch.float32, device=d.device) if C2 else None
            _chk(_L().fz_upsample2_concat_backward(_P(d.contiguous()), _P(da), _P(ds), Bq, Hq, Wq, C1, C2, _S()),
                 "fz_upsample2_concat_backward")
            if C2:
                dskips.append(ds)
            d = da
        return dskips + [d], grads      # skips were collected from block 2 (f4) down to block 0 (f16); d is now df32

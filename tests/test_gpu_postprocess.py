"""GPU parity: feeder + crop/argmax/softmax/accumulate kernels vs the numpy oracle (bit-exact
for argmax and the gathered bytes; |diff| <= 1 LSB allowed for round(softmax*255))."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(W, H, margin, P=128, res=0.2):
    from oracle.grid import Georef, generate_patches, tile_plan
    geo = Georef(700000.0, 6600000.0, res, W, H)
    tiles = generate_patches(P, margin, res, geo)
    return geo, tiles, tile_plan(tiles, geo, P, margin)


def _own(plan):
    from flair_for_aigle_b200.flair_zonal_detection.slicing import ownership_windows
    return ownership_windows(plan)


@pytest.mark.parametrize("W,H,margin", [(300, 217, 16), (256, 256, 32), (140, 411, 8)])
@pytest.mark.parametrize("layout,dtype", [("nchw", torch.float32), ("nhwc", torch.bfloat16), ("nhwc", torch.float32),
                                          ("nchw", torch.bfloat16)])
def test_crop_argmax_last_writer(cuda, W, H, margin, layout, dtype):
    from flair_for_aigle_b200 import native as nv
    from oracle.convert import write_tiles
    P, C = 128, 19
    geo, tiles, plan = _case(W, H, margin, P)
    n = len(tiles)
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((n, C, P, P)).astype(np.float32)
    logits = torch.from_numpy(logits).to(dtype).float().numpy()  # representable in dtype
    # force ties: np.argmax takes the first maximal index
    logits[:, 7, ::5, ::3] = logits.max(axis=1)[:, ::5, ::3]
    ref = np.full((H, W), 255, np.uint8)
    write_tiles(logits, plan, margin, ref, "argmax")
    t = torch.from_numpy(logits).to(cuda).to(dtype)
    if layout == "nhwc":
        cs = 24 if dtype == torch.bfloat16 else 20
        tt = torch.zeros((n, P, P, cs), dtype=dtype, device=cuda)
        tt[..., :C] = t.permute(0, 2, 3, 1)
        t = tt.contiguous()
    out = torch.full((H, W), 255, dtype=torch.uint8, device=cuda)
    plan_d = torch.from_numpy(plan).to(cuda)
    own_d = torch.from_numpy(_own(plan)).to(cuda)
    nv.crop_argmax_write(t, nv.NCHW if layout == "nchw" else nv.NHWC, margin, plan_d, own_d, out, n_cls=C)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)


def test_crop_softmax_write_and_accumulate(cuda):
    from flair_for_aigle_b200 import native as nv
    from oracle.convert import write_tiles, blend_accumulate, logits_to_labels_and_confidence
    P, C, margin = 128, 19, 16
    geo, tiles, plan = _case(300, 217, margin, P)
    n = len(tiles)
    rng = np.random.default_rng(6)
    logits = (rng.standard_normal((n, C, P, P)) * 3).astype(np.float32)
    ref = np.zeros((C, 217, 300), np.uint8)
    write_tiles(logits, plan, margin, ref, "class_prob")
    t = torch.from_numpy(logits).to(cuda)
    out = torch.zeros((C, 217, 300), dtype=torch.uint8, device=cuda)
    plan_d = torch.from_numpy(plan).to(cuda)
    own_d = torch.from_numpy(_own(plan)).to(cuda)
    nv.crop_softmax_write(t, nv.NCHW, margin, plan_d, own_d, out)
    torch.cuda.synchronize()
    diff = np.abs(out.cpu().numpy().astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3

    canvas_ref = np.zeros((C, 217, 300), np.float32)
    w = np.linspace(0.5, 1.5, (P - 2 * margin) ** 2, dtype=np.float32).reshape(P - 2 * margin, -1)
    blend_accumulate(logits, plan, margin, canvas_ref, w)
    canvas = torch.zeros((C, 217, 300), dtype=torch.float32, device=cuda)
    nv.crop_softmax_accumulate(t, nv.NCHW, margin, plan_d, torch.from_numpy(w).to(cuda), canvas)
    labels, conf = nv.canvas_argmax(canvas, want_confidence=True)
    torch.cuda.synchronize()
    assert np.abs(canvas.cpu().numpy() - canvas_ref).max() < 1e-5
    lab_ref, conf_ref = logits_to_labels_and_confidence(canvas.cpu().numpy())
    assert np.array_equal(labels.cpu().numpy(), lab_ref)
    assert np.array_equal(conf.cpu().numpy(), conf_ref)


@pytest.mark.parametrize("mode", ["argmax", "class_prob"])
def test_convert(cuda, mode):
    from flair_for_aigle_b200 import native as nv
    from oracle.convert import convert
    rng = np.random.default_rng(7)
    img = (rng.standard_normal((19, 61, 77)) * 2).astype(np.float32)
    img[3, ::2] = img.max(axis=0)[::2]
    ref = convert(img, mode)
    out = nv.convert(torch.from_numpy(img).to(cuda), 0 if mode == "argmax" else 1).cpu().numpy()
    if mode == "argmax":
        assert np.array_equal(out, ref)
    else:
        d = np.abs(out.astype(np.int16) - ref.astype(np.int16))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_feeder_matches_reference_read(cuda):
    from flair_for_aigle_b200 import native as nv
    from oracle.grid import read_tile
    P, margin = 128, 16
    geo, tiles, plan = _case(300, 217, margin, P)
    rng = np.random.default_rng(8)
    raster = rng.integers(0, 256, (4, 217, 300), dtype=np.uint8)
    mean = np.array([105.66, 111.35, 102.18, 106.59]); std = np.array([52.23, 45.62, 44.30, 39.78])
    ref = np.stack([read_tile(raster, int(r[0]), int(r[1]), P) for r in plan])
    refn = ((ref.astype(np.float64) - mean[None, :, None, None]) / std[None, :, None, None]).astype(np.float32)
    rd = torch.from_numpy(raster).to(cuda)
    org = torch.from_numpy(np.ascontiguousarray(plan[:, :2])).to(cuda)
    out = nv.gather_tiles_f32(rd, org, P, torch.tensor(mean, dtype=torch.float32, device=cuda),
                              torch.tensor(std, dtype=torch.float32, device=cuda))
    u8 = nv.gather_tiles_u8(rd, org, P)
    torch.cuda.synchronize()
    # mean/std are handed over as float32, the reference divides by the float64 config values
    assert np.abs(out.cpu().numpy() - refn).max() < 1e-6
    assert np.array_equal(u8.cpu().numpy(), ref.transpose(0, 2, 3, 1))

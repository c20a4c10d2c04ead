"""CPU: the conclusions of the precision budget (tests/error_budget.py) that DESIGN.md section 2 and csrc/operand.cuh rely
on, at a reduced tile size.  The full-size tables are profiles/r2_error_budget.txt (simulation) and
profiles/r2_stage_errors.txt (the engine against the simulation, stage by stage, on the B200)."""
import torch

import error_budget as eb


def test_fp16_operands_cut_the_rounding_noise_eightfold_and_the_simulator_is_the_oracle():
    torch.manual_seed(0)
    model = eb.make_model(1)
    x = eb.make_tile(101, 128)
    with torch.no_grad():
        ref = eb.simulate(model, x, {})
        want, _ = model({"AERIAL_RGBI": x, eb.TASK: torch.zeros(1, 19, 128, 128)})
        assert torch.allclose(ref, want[eb.TASK], rtol=0, atol=1e-4 * float(ref.std()))     # no rounding = the oracle
        sd = float(ref.std())
        err = {}
        for fmt in ("bf16", "fp16"):
            out = eb.simulate(model, x, {f: fmt for f in eb.FAMILIES})
            err[fmt] = float((out - ref).abs().mean()) / sd
            agree = float((out.argmax(1) == ref.argmax(1)).float().mean())
            print(f"all operand families {fmt}: mean|d|/std {err[fmt]:.5f}, class agreement {agree:.5f}")
        dec_only = float((eb.simulate(model, x, {"dec": "fp16"}) - ref).abs().mean()) / sd
        gelu = float((eb.simulate(model, x, {"gelu": "fit"}) - ref).abs().mean()) / sd
    assert 6.0 < err["bf16"] / err["fp16"] < 10.0            # 3 more significand bits
    assert err["fp16"] < 2e-3
    assert dec_only > 0.5 * err["fp16"]                      # the ten decoder activations dominate what is left
    assert gelu < 0.1 * err["fp16"]                          # the fitted GELU is far below the operand rounding


def test_engine_gelu_matches_erf_gelu():
    x = torch.linspace(-12, 12, 200001)
    d = (eb.gelu_engine(x, hw=False) - torch.nn.functional.gelu(x.double()).float()).abs().max().item()
    assert d < 3e-5


def test_what_lifts_the_floor_more_operand_bits_in_the_decoder():
    """profiles/r2_precision_floor.txt at a reduced size: with every family in fp16 (the engine), a hi / lo fp16 split of the
    decoder activations removes a third or more of the remaining logit error, the encoder's families much less, and a split
    of every family leaves nothing -- the lever DESIGN.md section 2b / 8 names for >= 99.9 % on every weight draw."""
    model = eb.make_model(1)
    x = eb.make_tile(101, 128)
    with torch.no_grad():
        ref = eb.simulate(model, x, {})
        sd = float(ref.std())
        base = {f: "fp16" for f in eb.FAMILIES}

        def err(fm):
            return float((eb.simulate(model, x, fm) - ref).abs().mean()) / sd
        e_all = err(base)
        e_dec_split = err({**base, "dec": "fp16x2"})
        e_dec_fp32 = err({**base, "dec": "fp32"})
        e_enc_fp32 = err({**base, **{f: "fp32" for f in ("y", "hidden", "w2s", "down")}})
        e_split = err({f: "fp16x2" for f in eb.FAMILIES})
    print(f"mean|d|/std: all fp16 {e_all:.5f}, dec hi/lo {e_dec_split:.5f}, dec fp32 {e_dec_fp32:.5f}, encoder fp32 {e_enc_fp32:.5f}, "
          f"all hi/lo {e_split:.6f}")
    assert e_dec_split < 0.75 * e_all and abs(e_dec_split - e_dec_fp32) < 0.02 * e_all      # 22 bits are as good as fp32 here
    assert e_enc_fp32 > e_dec_split                                                         # the decoder is the bigger lever
    assert e_split < 0.01 * e_all

"""Model goldens (tests/golden/model_*.npz, made by tests/golden/make_model_golden.py from the fp32 oracle):
* CPU: the oracle still reproduces them (pins the restatement against regressions);
* GPU: the sm_100a engines agree with them through the public model call, without the oracle at run time."""
import importlib.util
import os

import numpy as np
import pytest
import torch
from parity import (CLASS_AGREEMENT, CLASS_AGREEMENT_CONFIDENT, CLASS_AGREEMENT_FUSED, LOGIT_MAX_ABS,  # noqa: E402,F401
                    LOGIT_MEAN_ABS)

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_model_golden", os.path.join(HERE, "golden", "make_model_golden.py"))
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def _load(arch):
    return np.load(os.path.join(HERE, "golden", f"model_{arch.split('-')[0]}.npz"))


@pytest.mark.parametrize("arch", list(mg.ARCHS))
def test_oracle_reproduces_model_golden(arch):
    g = _load(arch)
    sd, _, xn = mg.golden_inputs(arch, mg.ARCHS[arch])
    s = mg.summarise(mg.oracle_logits(arch, sd, xn))
    # fp32 summation order depends on the host's thread count: allow it, but nothing more
    assert np.allclose(s["stats"], g["stats"], rtol=1e-4)
    assert np.abs(s["probe"] - g["probe"]).max() < 2e-3 * g["stats"][1]
    assert (s["classes"] == g["classes"]).mean() >= 0.9995
    conf = np.unpackbits(g["confident"]).astype(bool).reshape(256, 256)
    assert (s["classes"][conf] == g["classes"][conf]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("arch", list(mg.ARCHS))
def test_engine_matches_model_golden(cuda, arch):
    import bench
    from flair_for_aigle_b200.flair_hub.models.flair_model import FLAIR_HUB_Model
    from flair_for_aigle_b200.flair_zonal_detection.model_utils import prepare_model_config
    g = _load(arch)
    sd, _, xn = mg.golden_inputs(arch, mg.ARCHS[arch])
    c = bench.zonal_config("unused", "/tmp", "unused", 1)
    c["monotemp_arch"] = arch
    m = FLAIR_HUB_Model(prepare_model_config(c), {"AERIAL_RGBI": 512}, max_batch=1)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    out, _ = m({"AERIAL_RGBI": xn.to(cuda)})
    logits = out[mg.TASK][0].float().cpu()
    cls = logits[:, 128:384, 128:384].argmax(0).numpy().astype(np.uint8)
    conf = np.unpackbits(g["confident"]).astype(bool).reshape(256, 256)
    agree, agree_conf = (cls == g["classes"]).mean(), (cls[conf] == g["classes"][conf]).mean()
    probe_err = np.abs(logits[:, ::97, ::89].numpy() - g["probe"]).max() / g["stats"][1]
    print(f"{arch}: class agreement {agree:.5f}, on confident pixels {agree_conf:.6f}, probe max err / std {probe_err:.4f}")
    # bf16 operands end to end: stated tolerance as in tests/test_gpu_convnext.py::test_engine_vs_oracle
    assert agree >= CLASS_AGREEMENT and agree_conf >= CLASS_AGREEMENT_CONFIDENT and probe_err < LOGIT_MAX_ABS

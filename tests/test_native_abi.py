"""CPU: the C-ABI library builds, loads and exports every symbol include/flair_zonal_b200.h
declares; host-side argument validation works without a GPU; the product fails loudly (no CPU
fallback) when asked to compute without CUDA."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "flair_zonal_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fz_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from flair_for_aigle_b200.build import build_native
    from flair_for_aigle_b200 import native as nv
    lib_path = build_native()
    assert lib_path.exists()
    lib = ctypes.CDLL(str(lib_path))
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # the ctypes binding covers the same set
    assert sorted(nv.exported_symbols()) == declared
    assert nv.lib().fz_abi_version() == nv.ABI_VERSION == 3
    assert nv.lib().fz_operand_format() == nv.F16 and nv.op_dtype() == torch.float16     # the product build stores fp16


def test_argument_validation_without_gpu():
    """Shape checks run on the host before any launch."""
    from flair_for_aigle_b200 import native as nv
    lib = nv.lib()
    rc = lib.fz_gemm_bf16(None, None, None, None, None, None, 128, 100, 64, 1, 128, 0, None)   # N % 64 != 0
    assert rc == -1 and b"multiple of 64" in lib.fz_last_error()
    rc = lib.fz_gemm_bf16(None, None, None, None, None, None, 128, 128, 40, 1, 128, 0, None)    # K % 64 != 0
    assert rc == -1 and b"K=40" in lib.fz_last_error()
    rc = lib.fz_convert(None, 40, 4, 4, 0, None, None)                                           # > 32 classes
    assert rc == -1
    rc = lib.fz_convert(None, 19, 4, 4, 2, None, None)
    assert rc == -1 and b"Unknown output type" in lib.fz_last_error()
    rc = lib.fz_conv3x3_bf16(None, None, None, None, None, 1, 32, 32, 24, 16, 16, 0, 0, None, None, None, 0, 0, 0,
                             None)                                                               # Cin % 16 != 0
    assert rc == -1 and b"Cin=24" in lib.fz_last_error()
    # the training step's tile convolutions: pure host predicate + shape checks
    assert lib.fz_conv3x3_small_supported(512, 512, 32, 16) == 1 and lib.fz_conv3x3_small_supported(512, 512, 64, 32) == 1
    assert lib.fz_conv3x3_small_supported(510, 512, 32, 16) == 0            # H % 8
    assert lib.fz_conv3x3_small_supported(512, 500, 32, 16) == 0            # W % 32
    assert lib.fz_conv3x3_small_supported(512, 512, 128, 16) == 0           # more than 64 channels: the im2col + GEMM path
    rc = lib.fz_conv3x3_small_forward(None, None, None, None, 0, 1, 64, 64, 24, 16, 16, 16, 0, None)
    assert rc == -1 and b"not covered" in lib.fz_last_error()
    rc = lib.fz_conv3x3_small_wgrad(None, None, 64, None, 1, 64, 64, 32, 64, 0, None)          # Cout <= 32 only
    assert rc == -1 and b"Cout=64" in lib.fz_last_error()
    rc = lib.fz_confusion_matrix(None, None, 10, 200, None, None)                                # C <= 96
    assert rc == -1 and b"C=200" in lib.fz_last_error()
    rc = lib.fz_gelu_fwd_sumsq(None, None, None, None, 2, 64, 20, 1, None)                       # C % 8
    assert rc == -1 and b"C % 8" in lib.fz_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from flair_for_aigle_b200 import native as nv
    from flair_for_aigle_b200.flair_zonal_detection.postprocess import convert
    with pytest.raises(nv.NativeError):
        nv.gemm_bf16(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16), nv.EPI_BF16)
    with pytest.raises(nv.NativeError):
        convert(np.zeros((19, 4, 4), np.float32), "argmax")
    with pytest.raises(ValueError):
        convert(np.zeros((19, 4, 4), np.float32), "nope")          # same error as postprocess.py:29-30


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package or tools/ may import it (only tests/,
    __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs)."""
    for top in ("flair_for_aigle_b200", "tools"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dp, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)

"""Build libfz_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

``python -m flair_for_aigle_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU.  Objects are rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_native"
LIB = OUT_DIR / "libfz_b200.so"
INCLUDE = PKG.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_native(force: bool = False, verbose: bool = False, operands: str = "f16") -> Path:
    """operands="f16": the product library (inference kernels store fp16 operands, csrc/operand.cuh).
    operands="bf16": libfz_b200_bf16.so, the round-1 behaviour, for A/B measurements (FZ_OPERANDS=bf16 selects it)."""
    if operands not in ("f16", "bf16"):
        raise ValueError(operands)
    OUT_DIR.mkdir(exist_ok=True)
    obj_dir = OUT_DIR if operands == "f16" else OUT_DIR / "bf16"
    obj_dir.mkdir(exist_ok=True)
    lib = LIB if operands == "f16" else OUT_DIR / "libfz_b200_bf16.so"
    flags = NVCC_FLAGS + (["-DFZ_OPERANDS_BF16"] if operands == "bf16" else [])
    nvcc = _nvcc()
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [INCLUDE / "flair_zonal_b200.h"]
    objs = []
    jobs = []
    for src in sources:
        obj = obj_dir / (src.stem + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *flags, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = obj_dir / (src.stem + ".ptxas.log")
        log.write_text(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stderr[-4000:]}")
        if verbose:
            print(f"[build] {src.name} ok")
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or not lib.exists():
        cmd = [nvcc, "-shared", "-o", str(lib), *[str(o) for o in objs], "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stderr[-4000:]}")
        if verbose:
            print(f"[build] linked {lib}")
    return lib


RASTERIO_LIB = OUT_DIR / "libfz_rasterio.so"


def build_rasterio(force: bool = False, verbose: bool = False) -> Path:
    """libfz_rasterio.so: the host-only raster file I/O library (csrc/host/raster_io.cpp, include/flair_zonal_rasterio.h),
    g++ + zlib, no CUDA."""
    OUT_DIR.mkdir(exist_ok=True)
    src = CSRC / "host" / "raster_io.cpp"
    hdr = INCLUDE / "flair_zonal_rasterio.h"
    if force or _newer(RASTERIO_LIB, [src, hdr]):
        gxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
        cmd = [gxx, "-O3", "-std=c++17", "-Wall", "-fPIC", "-shared", "-pthread", "-I", str(INCLUDE), str(src), "-lz",
               "-o", str(RASTERIO_LIB)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"g++ failed for {src.name}:\n{res.stderr[-4000:]}")
        if verbose:
            print(f"[build] linked {RASTERIO_LIB}")
    return RASTERIO_LIB


if __name__ == "__main__":
    build_rasterio(force="--force" in sys.argv, verbose=True)
    build_native(force="--force" in sys.argv, verbose=True, operands="bf16" if "--bf16" in sys.argv else "f16")

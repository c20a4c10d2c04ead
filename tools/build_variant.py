"""Builds an A/B variant of the native library with extra nvcc flags: python tools/build_variant.py <name> -DFLAG [...]
-> flair_for_aigle_b200/_native/libfz_b200_<name>.so, loaded with FZ_LIB_VARIANT=<name> (objects under _native/<name>/)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flair_for_aigle_b200 import build as fb  # noqa: E402


def main() -> None:
    name, extra = sys.argv[1], sys.argv[2:]
    obj_dir = fb.OUT_DIR / name
    obj_dir.mkdir(parents=True, exist_ok=True)
    nvcc = fb._nvcc()
    sources = sorted(fb.CSRC.glob("*.cu"))

    def one(src):
        obj = obj_dir / (src.stem + ".o")
        res = subprocess.run([nvcc, *fb.NVCC_FLAGS, *extra, "-I", str(fb.INCLUDE), "-c", str(src), "-o", str(obj)],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"{src.name}:\n{res.stderr[-3000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, sources))
    lib = fb.OUT_DIR / f"libfz_b200_{name}.so"
    res = subprocess.run([nvcc, "-shared", "-o", str(lib), *[str(o) for o in objs], "-lcudart_static", "-ldl", "-lrt", "-lpthread"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stderr[-3000:])
    print(lib)


if __name__ == "__main__":
    main()

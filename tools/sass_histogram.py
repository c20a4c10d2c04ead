"""Opcode evidence from the BUILT objects (flair_for_aigle_b200/_native/*.o): per translation unit and per kernel, how many
tensor-core / TMEM / TMA / legacy-MMA instructions the sm_100a SASS holds.  Writes profiles/sass/<unit>.txt (committed:
the objects themselves are git-ignored) and profiles/sass/SUMMARY.txt.

    python tools/sass_histogram.py            (after `python -m flair_for_aigle_b200.build`; cuobjdump is in the CUDA toolkit)

What the mnemonics mean (B200_PROFILING.md): UTCHMMA = tcgen05.mma kind::f16 (UTCHMMA.2CTA = cta_group::2), LDTM = tcgen05.ld
(TMEM -> registers), UTMALDG = TMA load (cp.async.bulk.tensor global -> shared), UTMASTG = TMA store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, HMMA.16816 = legacy mma.sync, LDGSTS = cp.async, F2FP.SATFINITE.F16 = the saturating fp16 pack,
MUFU.EX2 / MUFU.RCP / MUFU.TANH = special-function unit ops.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "flair_for_aigle_b200", "_native")
OUT = os.path.join(ROOT, "profiles", "sass")
PAT = re.compile(r"\b(UTC[A-Z]*MMA\S*|UTCBAR\S*|UTMA[A-Z]+\S*|LDTM\S*|STTM\S*|HMMA\S*|IMMA\S*|LDGSTS\S*|SYNCS\S*|F2FP\S*|MUFU\.\S+|"
                 r"LDSM\S*|FFMA2?\b|FMUL2\b|FADD2\b|HFMA2\S*|REDUX\S*|ATOM\S*|RED\.\S*)")
KEY = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "F2FP.SATFINITE.F16", "F2FP.BF16", "MUFU.EX2",
       "MUFU.RCP", "MUFU.TANH", "FFMA")


def demangle(name: str) -> str:
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main() -> int:
    os.makedirs(OUT, exist_ok=True)
    summary = []
    for obj in sorted(f for f in os.listdir(NATIVE) if f.endswith(".o")):
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(NATIVE, obj)], capture_output=True, text=True)
        if sass.returncode != 0:
            print(f"cuobjdump failed on {obj}: {sass.stderr[-300:]}", file=sys.stderr)
            return 1
        per_kernel, cur = collections.OrderedDict(), None
        for line in sass.stdout.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = demangle(m.group(1))
                per_kernel[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and PAT.match(m.group(1)):
                per_kernel[cur][m.group(1)] += 1
        total = collections.Counter()
        lines = [f"# {obj}: opcode histogram of the sm_100a SASS (tools/sass_histogram.py)"]
        for k, c in per_kernel.items():
            total.update(c)
            if not c:
                continue
            short = re.sub(r"\(.*", "", k)
            lines.append(f"{short}")
            for op, n in sorted(c.items(), key=lambda kv: (-kv[1], kv[0])):
                lines.append(f"    {n:6d}  {op}")
        lines.insert(1, "TOTAL " + "  ".join(f"{op}={n}" for op, n in sorted(total.items(), key=lambda kv: (-kv[1], kv[0]))))
        open(os.path.join(OUT, obj[:-2] + ".txt"), "w").write("\n".join(lines) + "\n")

        def fam(prefix):
            return sum(n for op, n in total.items() if op.startswith(prefix))
        summary.append(f"{obj[:-2]:22s} kernels {len(per_kernel):3d}  " + "  ".join(
            f"{p}={fam(p)}" for p in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "F2FP.SATFINITE", "F2FP.BF16",
                                      "MUFU.EX2", "MUFU.TANH") if fam(p)))
    two_cta = subprocess.run("cuobjdump -sass %s | grep -c 'UTCHMMA.2CTA'" % os.path.join(NATIVE, "gemm_tcgen05_2sm.o"),
                             shell=True, capture_output=True, text=True).stdout.strip()
    summary.append(f"gemm_tcgen05_2sm: UTCHMMA.2CTA instructions = {two_cta}")
    open(os.path.join(OUT, "SUMMARY.txt"), "w").write(
        "# tensor-core / TMEM / TMA opcode counts per translation unit (whole object, all template instantiations)\n"
        + "\n".join(summary) + "\n")
    print("\n".join(summary))
    return 0


if __name__ == "__main__":
    sys.exit(main())

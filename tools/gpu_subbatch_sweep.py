"""Sweep the stage sub-batch sizes (L2 residency vs wave quantisation); eager total kernel time per tile."""
import os, sys, subprocess
for sb in ("0,0,0,0", "4,8,0,0", "4,8,18,0", "2,4,0,0", "8,16,0,0", "4,8,12,0", "4,4,0,0"):
    env = dict(os.environ, FZ_SUBBATCH=sb, B="37")
    out = subprocess.run([sys.executable, "tools/gpu_profile_batch.py"], env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if "ms/tile" in l]
    fam = [l for l in out.splitlines() if l.startswith(("gemm_tcgen05   ", "dwconv7_ln    ")) and "%" in l]
    print(sb, line[0] if line else out[-300:], " | ".join(f.strip() for f in fam))
